"""Generation + evaluation loop behind the reference's ``completion_eval.evaluate`` surface (reference:
pointnet2/completion_eval.py:65-330) for the dict-style datasets (mvp_dataset / shapenet_chunk / mvp40 / partnet).

Same keyword arguments, same per-batch arithmetic (network output / 2 / scale before the metrics), same return
tuple.  Differences, all on the host side:
  * ``parallel`` does not wrap anything in ``nn.DataParallel`` (completion_eval.py:113-118): the B200 design is one
    process per GPU, each iterating over its own contiguous shard of the dataset (dist.shard_range); with an
    initialised process group the per-shape metrics, labels and (if saved) clouds are all-gathered ONCE at the end;
  * the list-of-files ShapeNet loaders (``dataset='shapenet'``) need the reference's h5 directory walker and are
    not part of the hot path -> NotImplementedError;
  * with ``use_fused=True`` (default) the denoiser / refiner runs through the compiled sm_100a program.
"""
import os
import time

import numpy as np
import torch

from . import dist as pdist
from . import results_io
from .chamfer_loss_new import Chamfer_F1
from .emd import EMD_distance
from .point_upsample_module import point_upsample
from .util import sampling
from .util_fastdpmv2 import fast_sampling_function_v2

DICT_DATASETS = ("mvp_dataset", "shapenet_chunk", "mvp40", "partnet")


class AverageMeter(object):
    """util.py:7-26 without the tensorboard hook."""

    def __init__(self, name=""):
        self.name = name
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def generate_batch(net, data, diffusion_hyperparams, task="completion", refine_output_scale_factor=None,
                   save_multiple_t_slices=False, t_slices=(5, 10, 20, 50, 100, 200, 400, 600, 800),
                   use_a_precomputed_XT=False, T_step=100, point_upsample_factor=1,
                   include_displacement_center_to_final_output=False, noise_magnitude_added_to_gt=0.01,
                   add_noise_to_generated_for_refine_exp=False, fast_sampling=False, fast_sampling_config=None,
                   diffusion_config=None, print_every_n_steps=200, num_points=None, seed=None, noise_stream=0):
    """One batch of completion_eval.py:130-199 -> (generated (B,n,3) in network units, result_slices or None).
    ``seed`` None draws from the advancing process-wide noise stream (util._default_rng); with a seed, ``noise_stream``
    (batch index and rank, folded in by evaluate) keeps batches and ranks on different sub-streams."""
    dev = torch.device("cuda", torch.cuda.current_device())
    label = data["label"].to(dev)
    condition = data["partial"].to(dev)
    gt = data["complete"].to(dev)
    batch = gt.shape[0]
    num_points = gt.shape[1] if gt.dim() == 3 else num_points
    net.reset_cond_features()
    result_slices = None
    with torch.no_grad():
        if task == "refine_completion":
            generated = data["generated"].to(dev)
            if add_noise_to_generated_for_refine_exp:
                generated = generated + torch.normal(0, noise_magnitude_added_to_gt, size=generated.shape, device=dev)
            displacement = net(generated, condition, ts=None, label=label)
            if point_upsample_factor > 1:
                out, _ = point_upsample(generated, displacement, point_upsample_factor,
                                        include_displacement_center_to_final_output, refine_output_scale_factor)
            else:
                out = generated + displacement * refine_output_scale_factor
        elif task == "denoise":
            generated = gt + torch.normal(0, noise_magnitude_added_to_gt, size=gt.shape, device=dev)
            displacement = net(generated, condition=condition, ts=None, label=label)
            out = generated + displacement * refine_output_scale_factor
        else:
            XT = data["XT"].to(dev) if use_a_precomputed_XT else None
            size = (batch, num_points, 3)
            if save_multiple_t_slices:
                out, result_slices = sampling(net, size, diffusion_hyperparams, print_every_n_steps=print_every_n_steps,
                                              label=label, condition=condition, verbose=False,
                                              return_multiple_t_slices=True, t_slices=list(t_slices),
                                              use_a_precomputed_XT=use_a_precomputed_XT, step=T_step, XT=XT, seed=seed,
                                              noise_stream=noise_stream)
            elif fast_sampling:
                out = fast_sampling_function_v2(net, size, diffusion_hyperparams, diffusion_config,
                                                print_every_n_steps=print_every_n_steps, label=label, verbose=False,
                                                condition=condition, seed=seed, noise_stream=noise_stream,
                                                **(fast_sampling_config or {}))
            else:
                out = sampling(net, size, diffusion_hyperparams, print_every_n_steps=print_every_n_steps, label=label,
                               condition=condition, verbose=False, use_a_precomputed_XT=use_a_precomputed_XT,
                               step=T_step, XT=XT, seed=seed, noise_stream=noise_stream)
    return out, gt, label, result_slices


def evaluate(net, testloader, diffusion_hyperparams, print_every_n_steps=200, parallel=True, dataset="mvp_dataset",
             scale=1, save_generated_samples=False, save_dir=None, task="completion", refine_output_scale_factor=None,
             max_print_nums=1e8, save_multiple_t_slices=False, t_slices=(5, 10, 20, 50, 100, 200, 400, 600, 800),
             use_a_precomputed_XT=False, T_step=100, point_upsample_factor=1,
             include_displacement_center_to_final_output=False, compute_emd=True, compute_cd=True, num_points=None,
             augment_data_during_generation=False, noise_magnitude_added_to_gt=0.01,
             add_noise_to_generated_for_refine_exp=False, return_all_metrics=False, fast_sampling=False,
             fast_sampling_config=None, diffusion_config=None, use_fused=True, use_tf32=True, seed=None):
    """completion_eval.py:65-330.  ``testloader`` yields dicts with 'label', 'partial', 'complete' (+ 'generated' for
    task='refine_completion', 'XT' for use_a_precomputed_XT, 'M_inv'/'translation' for augmented generation)."""
    assert task in ["completion", "refine_completion", "denoise"]
    if fast_sampling:
        assert not save_multiple_t_slices
        assert not use_a_precomputed_XT
    if dataset in ("shapenet", "shapenet_pytorch"):
        raise NotImplementedError("the list-of-files ShapeNet loaders are outside the hot path; use a dict-style dataset")
    if dataset not in DICT_DATASETS:
        raise Exception("%s dataset is not supported" % dataset)
    if use_a_precomputed_XT or augment_data_during_generation:
        assert task == "completion"
    CD_meter, F1_meter, EMD_meter = AverageMeter(), AverageMeter(), AverageMeter()
    dev = torch.device("cuda", torch.cuda.current_device())
    total_meta = torch.zeros(0, dtype=torch.long, device=dev)
    metrics = {k: torch.zeros(0, device=dev) for k in ("cd_distance", "emd_distance", "cd_p", "f1")}
    f1_threshold = 0.001 if dataset == "mvp40" else 0.0001          # :103
    cd_module, emd_module = Chamfer_F1(f1_threshold=f1_threshold), EMD_distance()
    if use_fused and hasattr(net, "enable_fused"):
        net.enable_fused(True, use_tf32=use_tf32, use_graph=True, fuse_cold=True)
    total_len = len(testloader)
    print_interval = int(np.ceil(total_len / max_print_nums))
    total_time = 0
    kept, kept_slices = [], {}
    for idx, data in enumerate(testloader):
        start = time.time()
        generated_data, gt, label, result_slices = generate_batch(
            net, data, diffusion_hyperparams, task=task, refine_output_scale_factor=refine_output_scale_factor,
            save_multiple_t_slices=save_multiple_t_slices, t_slices=t_slices, use_a_precomputed_XT=use_a_precomputed_XT,
            T_step=T_step, point_upsample_factor=point_upsample_factor,
            include_displacement_center_to_final_output=include_displacement_center_to_final_output,
            noise_magnitude_added_to_gt=noise_magnitude_added_to_gt,
            add_noise_to_generated_for_refine_exp=add_noise_to_generated_for_refine_exp, fast_sampling=fast_sampling,
            fast_sampling_config=fast_sampling_config, diffusion_config=diffusion_config,
            print_every_n_steps=print_every_n_steps, num_points=num_points, seed=seed,
            noise_stream=1 + idx + total_len * pdist.rank())       # one noise sub-stream per (batch, rank)
        torch.cuda.synchronize()
        generation_time = time.time() - start
        total_time += generation_time
        batch = gt.shape[0]
        if augment_data_during_generation:                            # :203-205
            M_inv, translation = data["M_inv"].to(dev), data["translation"].to(dev)
            generated_data = torch.matmul(generated_data - translation, M_inv)
            gt = torch.matmul(gt - translation, M_inv)
        generated_data = generated_data / 2 / scale                    # :206-207
        gt = gt / 2 / scale
        if result_slices is not None:
            for key in result_slices:
                v = result_slices[key]
                if augment_data_during_generation:
                    v = torch.matmul(v - translation, M_inv)
                result_slices[key] = (v / 2 / scale).detach()
        if compute_cd:                                                 # :217-227
            cd_p, dist, f1 = cd_module(generated_data, gt)
        else:
            dist = torch.zeros(batch, device=dev)
            cd_p = f1 = dist
        emd_cost = emd_module(generated_data, gt) if compute_emd else torch.zeros_like(dist)
        total_meta = torch.cat([total_meta, label])
        for k, v in (("cd_distance", dist), ("emd_distance", emd_cost), ("cd_p", cd_p), ("f1", f1)):
            metrics[k] = torch.cat([metrics[k], v])
        CD_meter.update(dist.mean().item(), n=batch)
        F1_meter.update(f1.mean().item(), n=batch)
        EMD_meter.update(emd_cost.mean().item(), n=batch)
        if idx % print_interval == 0:
            print("progress [%d/%d] %.4f (%d samples) CD distance %.8f EMD distance %.8f F1 score %.6f this batch time "
                  "%.2f total generation time %.2f" % (idx, total_len, idx / total_len, batch, CD_meter.avg,
                                                       EMD_meter.avg, F1_meter.avg, generation_time, total_time),
                  flush=True)
        if save_generated_samples:
            kept.append(generated_data)
            if result_slices is not None:
                for t, v in result_slices.items():
                    kept_slices.setdefault(t, []).append(v)
    # ---- one gather at the end (replaces DataParallel per step and the file-system gather of the drivers) --------
    if parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
        total_meta = pdist.all_gather_shapes(total_meta)
        metrics = {k: pdist.all_gather_shapes(v) for k, v in metrics.items()}
        if save_generated_samples and kept:
            kept = [pdist.all_gather_shapes(torch.cat(kept, 0))]
            # the T-slice clouds go through the same gather, so every *_T<t>.h5 holds as many rows as the main file
            kept_slices = {t: [pdist.all_gather_shapes(torch.cat(parts, 0))] for t, parts in sorted(kept_slices.items())}
        rank = torch.distributed.get_rank()
        avg_cd = metrics["cd_distance"].mean().item() if metrics["cd_distance"].numel() else 0.0
        avg_emd = metrics["emd_distance"].mean().item() if metrics["emd_distance"].numel() else 0.0
    else:
        rank = 0
        avg_cd, avg_emd = CD_meter.avg, EMD_meter.avg                # :327-330
    if save_generated_samples and kept and rank == 0:
        os.makedirs(save_dir, exist_ok=True)
        n_pts = kept[0].shape[1]
        results_io.save_generated(os.path.join(save_dir, results_io.generated_file_name(dataset, n_pts)),
                                  torch.cat(kept, 0).detach().cpu().numpy())
        for t, parts in kept_slices.items():
            results_io.save_generated(os.path.join(save_dir, results_io.generated_file_name(dataset, n_pts, t)),
                                      torch.cat(parts, 0).detach().cpu().numpy())
    total_meta = total_meta.detach().cpu().numpy()
    if return_all_metrics:
        return avg_cd, avg_emd, total_meta, metrics
    return avg_cd, avg_emd, total_meta, metrics["cd_distance"], metrics["emd_distance"]
