#!/usr/bin/env python
"""bench.py -- throughput of the hot path on B200, one JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port), rank 0 only

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): conditional DDPM reverse
sampling, B = 32 shapes per GPU, 2048-point output, 3072-point mirrored partial condition (4 channels),
random-init PointNet2CloudCondition (9.76 M parameters, shipped DDPM config), T = 1000.
A *step* is one pass of the hot path over one batch: eps_theta(x_t, t, c) with retained condition
features + the posterior update with fresh device noise.  metric = shapes/s of the full T=1000 chain
= (world * B) / (T * seconds per step); steps are warm steps (the one cold step per chain, which also
encodes the condition cloud, is timed separately and reported as `cold_ms`).
`e2e` runs util.sampling() -- the public API -- from pinned HOST buffers (condition, labels, x_T copied
H2D, result copied D2H inside the timed region) for `--e2e-steps` reverse steps, cold step included.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
JSON_OUT = [sys.stdout]      # where the single JSON line goes; main() points sys.stdout at stderr for everything else
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_CHAIN = 1000
CPU_BASELINE_STEPS = 8        # warm steps of the cpu_baseline leg of the main arm: ~10 s of host work at 4 shapes/step
N_POINTS, M_COND = 2048, 3072
WORKLOAD = "ddpm_reverse_step B=32/gpu N=2048 cond=3072x4 T=1000 PointNet2CloudCondition(9.76M, random init)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=32, help="shapes per GPU")
    p.add_argument("--e2e-steps", type=int, default=200)
    p.add_argument("--cpu-batch", type=int, default=4, help="shapes per step of the CPU arm (bounded sample)")
    p.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arm (0 = min(cores, 32))")
    p.add_argument("--no-tf32", action="store_true", help="fp32 SIMT GEMMs instead of TF32 tensor cores")
    p.add_argument("--engine", default="fused", choices=["fused", "modules"],
                   help="fused: compiled program of our kernels (fused.py); modules: per-layer torch modules")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--module-cold", action="store_true",
                   help="encode the condition cloud with the per-layer torch modules instead of the compiled condition program")
    p.add_argument("--dump-ops", default="", help="write the per-op timing of one eager program replay to this JSON file")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-eval-kernels", action="store_true", help="skip the Chamfer / EMD Mpairs/s lines (configs[3])")
    p.add_argument("--no-fast-ddpm", action="store_true",
                   help="skip configs[4]: FastDPM 50-step VAR chain at B=128/GPU (+ on rank 0 its CD against T=1000 chains)")
    p.add_argument("--no-gpu-reference", action="store_true",
                   help="skip the `gpu_reference` leg (reference CUDA kernels + cuDNN module path) and the fp32-SIMT leg")
    p.add_argument("--no-strong", action="store_true", help="skip the strong-scaling line (32 shapes total) at N > 1")
    p.add_argument("--profiler-range", action="store_true",
                   help="bracket the timed region with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic inputs and model
# ------------------------------------------------------------------------------------------------------
def make_inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    uvw = torch.rand(B, M_COND, 3, generator=g) * 2 - 1
    flag = torch.ones(B, M_COND, 1)
    flag[:, M_COND // 2:] = -1
    cond = torch.cat([uvw, flag], dim=2)
    label = torch.arange(B) % 16
    xT = torch.randn(B, N_POINTS, 3, generator=g)
    return cond, label, xT


def make_net(device):
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    torch.manual_seed(0)
    return PointNet2CloudCondition(configs.ddpm_pointnet_config()).eval().to(device)


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region (profiling recipe)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# per-kernel accounting: which of OUR kernels dominates the step, and its roofline
# ------------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, a):
    """Algorithmic HBM bytes of one C-ABI call from its integer arguments (per-kernel table)."""
    def num(v):
        v = getattr(v, "value", v)
        return int(v) if isinstance(v, (int, float)) else 0
    i = [num(v) for v in a[:6]]
    if name == "pdr_furthest_point_sampling":
        b, n, m = i[:3]
        return b * (12 * n + 4 * m)
    if name == "pdr_ball_query":
        b, n, m = i[:3]
        ns = num(a[4])
        return b * (12 * (n + m) + 4 * m * (ns + 1))
    if name == "pdr_group_points":
        b, c, n, npt, ns = i[:5]
        return b * (4 * npt * ns + 4 * c * npt * ns + 4 * c * min(n, npt * ns))
    if name == "pdr_gather_points":
        b, c, n, m = i[:4]
        return b * (4 * m + 8 * c * m)
    if name == "pdr_knn_points":
        b, p1, p2, k = i[:4]
        return b * (12 * (p1 + p2) + 12 * p1 * k)
    if name == "pdr_affine_noise_update":
        return 12 * i[0]
    return None


class KernelAccounting:
    """Wraps point_diffusion_refinement_b200._lib.call with CUDA events on the launching stream."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        from point_diffusion_refinement_b200 import _lib
        self._lib, self._orig = _lib, _lib.call
        mods = [m for n, m in sys.modules.items() if n.startswith("point_diffusion_refinement_b200") and hasattr(m, "call")]
        self._mods = mods

        def wrapped(name, *args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self._orig(name, *args)
            e.record()
            self.records.append((name, args, s, e))

        for m in mods:
            m.call = wrapped
        return self

    def __exit__(self, *exc):
        for m in self._mods:
            m.call = self._orig

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, args, s, e in self.records:
            ms = s.elapsed_time(e)
            d = agg.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0})
            d["calls"] += 1
            d["ms"] += ms
            b = algorithmic_bytes(name, args)
            d["bytes"] += b or 0
        return agg


# ------------------------------------------------------------------------------------------------------
# CPU arm: the same step through the package's host modules bound to the CPU oracle
# ------------------------------------------------------------------------------------------------------
def cpu_threads(requested=0):
    """Threads for the CPU arm.  Above ~32 threads the many small per-layer CPU kernels of this workload only
    contend (measured: 128 threads are 10x SLOWER than 8 on the same step), so the arm uses min(cores, 32)."""
    n = requested if requested > 0 else min(os.cpu_count() or 1, 32)
    os.environ["OMP_NUM_THREADS"] = str(n)      # before the oracle's libgomp is loaded
    torch.set_num_threads(n)
    return n


def cpu_step_seconds(batch, steps, warmup):
    """Seconds per warm DDPM step on the host cores (oracle port: C/OpenMP ops + CPU torch for the MLPs)."""
    from tests import common as C
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    torch.manual_seed(0)
    net = PointNet2CloudCondition(configs.ddpm_pointnet_config()).eval()
    cond, label, x = make_inputs(batch, seed=1)
    ts = torch.full((batch,), 999.0)
    with C.package_bound_to_oracle(), torch.no_grad():
        eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)  # cold
        for _ in range(max(warmup - 1, 0)):
            eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
            x = (x - 0.01 * eps) * 1.0001 + 0.01 * torch.randn_like(x)
        dt = (time.perf_counter() - t0) / steps
    return dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = cpu_threads(args.cpu_threads)
    dt = cpu_step_seconds(args.cpu_batch, args.steps, args.warmup)
    value = args.cpu_batch / (T_CHAIN * dt)
    sample = "%d shapes/step x %d warm steps of the same denoise step (oracle C/OpenMP ops + CPU torch MLPs)" % (
        args.cpu_batch, args.steps)
    line = {"impl": "reference", "metric": "ddpm_shapes_per_sec_T1000", "value": value, "unit": "shapes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "the reference has no CPU implementation of these ops "
                       "(AT_ASSERT CPU not supported); this arm is the oracle port on the host cores",
                       "scope": "ONE host CPU whatever --gpus is (rank 0 only): this value does not grow with N, so a "
                                "ratio against an N-GPU value is N-inflated -- compare at n_gpus = 1",
                       "n_hosts": 1},
            "cpu_baseline": {"value": value, "unit": "shapes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True, file=JSON_OUT[0])


# ------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: Chamfer / EMD evaluation kernels, Mpairs/s (SURVEY 8d cfg 4)
# ------------------------------------------------------------------------------------------------------
def eval_kernel_lines(dev, sm_mhz):
    """Chamfer_F1 and EMD_distance through the reference-facing modules, B=256.  1 pair = one (i,j) squared
    distance of one cloud pair, counted once.  Both kernels are FP32-ALU/MUFU bound, not HBM bound
    (inputs are 6-100 MB, L2 or not is irrelevant), so the fraction quoted is of the FP32 lane rate
    148 SM x 128 lanes x clk: Chamfer evaluates each pair in both directions (2 evals/pair, ~7 instr each),
    EMD runs 30 exp-weighted passes per pair."""
    from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1
    from point_diffusion_refinement_b200.emd import EMD_distance
    g = torch.Generator().manual_seed(7)
    lanes_per_s = 148 * 128 * (sm_mhz or 1965) * 1e6
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for name, Bq, n, reps in (("chamfer_f1_256x2048x2048", 256, 2048, 10), ("chamfer_f1_256x16384x16384", 256, 16384, 3)):
        a = (torch.rand(Bq, n, 3, generator=g) * 2 - 1).to(dev); b = (torch.rand(Bq, n, 3, generator=g) * 2 - 1).to(dev)
        cf = Chamfer_F1()
        ms = timed(lambda: cf(a, b), reps)
        pairs = Bq * n * n
        out[name] = {"ms": ms, "mpairs_per_s": pairs / ms / 1e3, "algorithmic_GBps": Bq * 16 * 2 * n / ms / 1e6,
                     "pair_evals_per_lane_clk": 2 * pairs / (ms * 1e-3) / lanes_per_s}
        del a, b
    a = (torch.rand(256, 2048, 3, generator=g)).to(dev); b = (torch.rand(256, 2048, 3, generator=g)).to(dev)
    em = EMD_distance()
    ms = timed(lambda: em(a, b), 3)
    pairs = 256 * 2048 * 2048
    out["emd_256x2048x2048"] = {"ms": ms, "mpairs_per_s": pairs / ms / 1e3,
                                "exp_pair_evals_per_lane_clk": 30 * pairs / (ms * 1e-3) / lanes_per_s}
    return out


def geometry_kernel_lines(dev, sm_mhz):
    """BASELINE configs[0] (SURVEY 8d cfg 1): FPS 4096 -> 1024 on one cloud + ball query of the picks, and the same two
    kernels at the sampler's batch.  FPS is a serial chain of m-1 argmax rounds (latency-bound): quoted as distance
    evaluations per FP32 lane-clock; ball query as centre-point tests per lane-clock (early exit makes it an upper
    bound on work)."""
    from point_diffusion_refinement_b200 import _ext
    g = torch.Generator().manual_seed(0)
    lanes_per_s = 148 * 128 * (sm_mhz or 1965) * 1e6
    out = {}
    for name, Bq, n, m, reps in (("fps_1x4096_to_1024", 1, 4096, 1024, 20), ("fps_32x2048_to_1024", 32, 2048, 1024, 20)):
        xyz = (torch.rand(Bq, n, 3, generator=g) * 2 - 1).to(dev)
        for _ in range(3):
            idx = _ext.furthest_point_sampling(xyz, m)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            idx = _ext.furthest_point_sampling(xyz, m)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        evals = Bq * (m - 1) * n
        out[name] = {"ms": ms, "us_per_round": ms * 1e3 / (m - 1), "dist_evals_per_s": evals / (ms * 1e-3),
                     "dist_evals_per_lane_clk": evals / (ms * 1e-3) / lanes_per_s,
                     "algorithmic_GBps": Bq * (12 * n + 4 * m) / ms / 1e6}
        centres = _ext.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        for _ in range(3):
            _ext.ball_query(centres, xyz, 0.2, 32)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            _ext.ball_query(centres, xyz, 0.2, 32)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name.replace("fps", "ball_query_r0.2_ns32").replace("_to_", "_x_")] = {
            "ms": ms, "tests_per_lane_clk_upper": Bq * m * n / (ms * 1e-3) / lanes_per_s,
            "algorithmic_GBps": Bq * (12 * (n + m) + 4 * m * 33) / ms / 1e6}
    return out


def fast_ddpm_lines(net, dev, B, dh, cond, label, rank, world, with_cd):
    """BASELINE configs[4]: fast_sampling_function_v2(length=50, 'var', 'quadratic', kappa=0.5) (README.md:95).
    Throughput: 128 shapes per GPU (the configuration's batch), every rank, max over ranks.  Quality (rank 0, the bench
    batch): cd_t between the 50-step output and a full T=1000 chain for the same condition clouds, with the cd_t between
    two T=1000 seeds as the noise floor of a random-init network."""
    import contextlib
    import io
    import torch.distributed as dist
    from point_diffusion_refinement_b200 import util, util_fastdpmv2
    from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1
    dcfg = {"T": T_CHAIN, "beta_0": 1e-4, "beta_T": 0.02}

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with contextlib.redirect_stdout(io.StringIO()):
            e0.record(); r = fn(); e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return r, ms.item() / 1e3

    def fast(c, l, seed):
        return util_fastdpmv2.fast_sampling_function_v2(
            net, (c.shape[0], N_POINTS, 3), dh, dcfg, length=50, sampling_method="var", schedule="quadratic", kappa=0.5,
            print_every_n_steps=0, label=l, verbose=False, condition=c, seed=seed, noise_stream=rank)

    B5 = 128
    cond5, label5, _ = make_inputs(B5, seed=500 + rank)
    cond5, label5 = cond5.to(dev), label5.to(dev)
    with contextlib.redirect_stdout(io.StringIO()):
        fast(cond5, label5, 0)                                  # compiles the B=128 programs, warms the API path
    _, t_fast = timed(lambda: fast(cond5, label5, 1))
    out = {"fast50_var_quadratic_kappa0.5": {"batch_per_gpu": B5, "seconds": t_fast, "net_calls": 50,
                                             "shapes_per_s": world * B5 / t_fast, "n_gpus": world,
                                             "ms_per_net_call": t_fast * 1e3 / 50}}
    del cond5, label5
    net._fused_engine = None                                    # drop the B=128 buffers
    if with_cd:
        size = (B, N_POINTS, 3)
        full = lambda seed: util.sampling(net, size, dh, print_every_n_steps=0, label=label, verbose=False,
                                          condition=cond, seed=seed)
        with contextlib.redirect_stdout(io.StringIO()):
            xf = fast(cond, label, 1)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); x1 = full(1); torch.cuda.synchronize(); t_full = time.perf_counter() - t0
            x2 = full(1001)
        cf = Chamfer_F1()
        cd = lambda a, b: float(cf(a / 2, b / 2)[1].mean().item())
        out.update({"ddpm_T1000": {"batch_per_gpu": B, "seconds": t_full, "shapes_per_s_per_gpu": B / t_full,
                                   "net_calls": 1000},
                    "cd_t_fast_vs_T1000": cd(xf, x1), "cd_t_T1000_seed_vs_seed": cd(x1, x2),
                    "note": "random-init network: the CD values only show that the 50-step chain lands as close to a "
                            "T=1000 sample as another T=1000 sample does"})
    return out


def timed_steps(step_fn, steps, warmup, dev, world):
    """`steps` calls of step_fn timed with CUDA events, barrier + device sync on both sides, max over ranks -> ms/step."""
    import torch.distributed as dist

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(warmup):
        step_fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / steps


def baseline_legs(net, dev, B, dh, cond, label, xT_h, steps):
    """N = 1 only.  (1) `fp32_simt`: the same compiled step with fp32 SIMT GEMMs instead of TF32 tensor cores (the
    fp32-faithful number).  (2) `gpu_reference`: the reference's DESIGN on this box -- per-layer torch modules (cuDNN 1x1
    convolutions, ATen GroupNorm) with the grouping / sampling entry points bound to the reference's own CUDA kernels
    recompiled for sm_100a (oracle/_ref/libpdr_ref_cuda.so; pointnet2_with_pcld_condition.py:276-476 via util.py:224-249);
    TF32 convolutions as PyTorch defaults to on this GPU.  Baselines, not the product path."""
    from point_diffusion_refinement_b200 import util
    out = {}
    rng = util.DeviceNoise(seed=77)
    Alpha = dh["Alpha"].numpy(); Abar = dh["Alpha_bar"].numpy(); Sigma = dh["Sigma"].numpy()
    ts = torch.empty((B,), dtype=torch.float32, device=dev)
    x = xT_h.to(dev)
    state = {"t": T_CHAIN - 1}

    def one_step():
        t = state["t"]; state["t"] -= 1
        ts.fill_(float(t))
        eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
        inv = 1.0 / float(Alpha[t]) ** 0.5
        c = (1.0 - float(Alpha[t])) / (1.0 - float(Abar[t])) ** 0.5
        rng.affine_update(x, eps.contiguous(), inv, -c * inv, float(Sigma[t]))

    net.reset_cond_features()
    net.enable_fused(True, use_tf32=False, use_graph=True, fuse_cold=True)
    one_step()                                                   # cold
    out["fp32_simt"] = {"ms_per_step": timed_steps(one_step, steps, 3, dev, 1), "steps": steps,
                        "engine": "fused program, fp32 SIMT GEMMs (use_tf32=False)"}
    out["fp32_simt"]["shapes_per_s"] = B / (T_CHAIN * out["fp32_simt"]["ms_per_step"] / 1e3)
    net.reset_cond_features()
    net.enable_fused(False)
    try:
        from oracle import ref_cuda
        from tests import common as C
        if not ref_cuda.available():
            raise RuntimeError("oracle/_ref/libpdr_ref_cuda.so not present")
        old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = True
        x.copy_(xT_h); state["t"] = T_CHAIN - 1
        try:
            with C.package_bound_to_reference_cuda():
                one_step()                                       # cold: encodes the condition cloud
                ms = timed_steps(one_step, steps, 2, dev, 1)
        finally:
            torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        out["gpu_reference"] = {"ms_per_step": ms, "shapes_per_s": B / (T_CHAIN * ms / 1e3), "steps": steps, "kind": "reference-design",
                                "what": "per-layer torch modules (cuDNN convs, TF32 allowed as by PyTorch default) + the reference's own "
                                        "pointnet2_ops CUDA kernels recompiled for sm_100a (oracle/_ref); kNN on our kernel "
                                        "(pytorch3d is not vendored by the reference)"}
    except Exception as exc:       # the .so is built where /root/reference exists and travels with the snapshot
        out["gpu_reference"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    net.reset_cond_features()
    return out


# ------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    # contract: stdout carries exactly ONE JSON line.  Everything the library prints on the way (the reference's
    # "begin sampling ..." banners, kept for drop-in fidelity) goes to stderr.
    json_out = sys.stdout
    sys.stdout = sys.stderr
    try:
        _main(args, json_out)
    finally:
        sys.stdout = json_out


def _main(args, json_out):
    JSON_OUT[0] = json_out
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from point_diffusion_refinement_b200 import _lib, util
    from point_diffusion_refinement_b200 import dist as pdist
    _lib.lib()  # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        pdist.init_from_env(backend="nccl")
    tf32 = not args.no_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True

    B = args.batch
    net = make_net(dev)
    cond_h, label_h, xT_h = [t.pin_memory() for t in make_inputs(B, seed=100 + rank)]
    cond, label, x = cond_h.to(dev), label_h.to(dev), xT_h.to(dev)
    dh = util.calc_diffusion_hyperparams(T=T_CHAIN, beta_0=1e-4, beta_T=0.02)
    rng = util.DeviceNoise(seed=1234 + rank)
    ts = torch.empty((B,), dtype=torch.float32, device=dev)
    Alpha = dh["Alpha"].numpy(); Abar = dh["Alpha_bar"].numpy(); Sigma = dh["Sigma"].numpy()

    def one_step(t):
        nonlocal x
        ts.fill_(float(t))
        eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
        inv = 1.0 / float(Alpha[t]) ** 0.5
        c = (1.0 - float(Alpha[t])) / (1.0 - float(Abar[t])) ** 0.5
        rng.affine_update(x, eps.contiguous(), inv, -c * inv, float(Sigma[t]))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        # cold step (encodes the condition cloud), timed on its own
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fuse_cold = args.engine == "fused" and not args.module_cold
        if args.engine == "fused":
            # tcgen05 TF32 GEMMs (gemm_tc.cu); fuse_cold: the condition branch runs as a compiled program too
            net.enable_fused(True, use_tf32=tf32, use_graph=not args.no_graph, fuse_cold=fuse_cold)
        e0.record(); one_step(T_CHAIN - 1); e1.record(); torch.cuda.synchronize()
        first_call_ms = e0.elapsed_time(e1)            # includes building / capturing the programs (once per process)
        net.reset_cond_features()
        x.copy_(xT_h, non_blocking=True)
        sync()
        e0.record(); one_step(T_CHAIN - 1); e1.record(); torch.cuda.synchronize()
        cold_ms = e0.elapsed_time(e1)                  # a cold step of a later chain: condition branch + x branch
        t = T_CHAIN - 2
        for _ in range(max(args.warmup, 3)):
            one_step(t); t -= 1
        # ---- timed region: exactly K warm steps ------------------------------------------------------
        clocks = ClockSampler(local_rank)
        sync()
        clocks.start()
        launches0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        for _ in range(args.steps):
            one_step(t); t -= 1
        e1.record()
        sync()
        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStop()
        clock_info = clocks.stop()
        launches = _lib.launch_count - launches0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_per_step = ms.item() / args.steps

        # ---- which of our kernels dominates, and its roofline (one extra, untimed, instrumented step) --
        eng = getattr(net, "_fused_engine", None)
        if eng is not None:
            agg = eng.profile()
            agg.pop("torch", None)
            if args.dump_ops and rank == 0:
                with open(args.dump_ops, "w") as f:
                    json.dump(eng.last_profile, f)
            launches = args.steps * (eng.n_kernel_calls + 1)      # program kernels + the fused update, per step
        else:
            with KernelAccounting() as acct:
                one_step(t); t -= 1
                agg = acct.summary()
        net.reset_cond_features()

        # ---- e2e through the public API with host buffers ----------------------------------------------
        e2e = None
        if not args.no_e2e:
            Ke = min(args.e2e_steps, T_CHAIN - 1)
            out_h = torch.empty((world * B, N_POINTS, 3), dtype=torch.float32).pin_memory() if rank == 0 else None

            def chain():
                c = cond_h.to(dev, non_blocking=True); l = label_h.to(dev, non_blocking=True)
                xT = xT_h.to(dev, non_blocking=True)
                r = util.sampling(net, (B, N_POINTS, 3), dh, label=l, condition=c, verbose=False,
                                  print_every_n_steps=0, use_a_precomputed_XT=True, step=Ke, XT=xT, seed=7, noise_stream=rank)
                r = pdist.all_gather_shapes(r)                 # the one collective of the path (final gather), timed
                if rank == 0:
                    out_h.copy_(r, non_blocking=True)          # the whole job's clouds land on the host of rank 0
                return r

            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                chain()                       # warm-up of the API path
                sync()
                t0 = time.perf_counter()
                e0.record(); r = chain(); e1.record()
                sync()
                wall = time.perf_counter() - t0
            ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            chain_s = ems.item() / 1e3
            h2d = world * (cond_h.numel() * 4 + label_h.numel() * 8 + xT_h.numel() * 4)
            e2e = {"value": world * B / (chain_s * T_CHAIN / Ke), "unit": "shapes/s",
                   "h2d_bytes_per_step": h2d / Ke, "d2h_bytes_per_step": world * B * N_POINTS * 3 * 4 / Ke,
                   "chain_steps": Ke, "chain_ms": chain_s * 1e3, "cold_steps_in_chain": 1,
                   "note": "util.sampling(use_a_precomputed_XT, step=%d): %d reverse steps incl. the cold one (= %.1f warm "
                           "steps), pinned-host condition/label/x_T in, NCCL all_gather of the clouds + D2H of the whole "
                           "job's result on rank 0 inside the timed region; scaled by T/steps (which charges the cold step "
                           "and the gather %dx per chain)" % (Ke, Ke, cold_ms / ms_per_step, T_CHAIN // Ke)}

    # ---- strong scaling (SURVEY 8d cfg 3, secondary line): 32 shapes in total, 32 / N per GPU ---------------------
    strong = None
    if world > 1 and not args.no_strong and args.engine == "fused" and 32 % world == 0:
        with torch.no_grad():
            Bs = 32 // world
            cond_s, label_s, x_s = [v.to(dev) for v in make_inputs(Bs, seed=300 + rank)]
            ts_s = torch.empty((Bs,), dtype=torch.float32, device=dev)
            st = {"t": T_CHAIN - 1}

            def strong_step():
                tt = st["t"]; st["t"] -= 1
                ts_s.fill_(float(tt))
                eps = net(x_s, cond_s, ts=ts_s, label=label_s, use_retained_condition_feature=True)
                inv = 1.0 / float(Alpha[tt]) ** 0.5
                c = (1.0 - float(Alpha[tt])) / (1.0 - float(Abar[tt])) ** 0.5
                rng.affine_update(x_s, eps.contiguous(), inv, -c * inv, float(Sigma[tt]))

            net.reset_cond_features()
            strong_step()                                           # cold; compiles the (32 / N)-shape programs
            s_ms = timed_steps(strong_step, args.steps, max(args.warmup, 3), dev, world)
            eng_s = getattr(net, "_fused_engine", None)
            per_kernel = None
            if eng_s is not None and rank == 0:
                agg_s = eng_s.profile(); agg_s.pop("torch", None)
                per_kernel = {k: round(v["ms"], 4) for k, v in sorted(agg_s.items(), key=lambda kv: -kv[1]["ms"])}
            net.reset_cond_features()
            strong = {"scaling": "strong", "batch_total": 32, "batch_per_gpu": Bs, "n_gpus": world, "ms_per_step": s_ms,
                      "shapes_per_s": 32 / (T_CHAIN * s_ms / 1e3), "per_kernel_ms_eager": per_kernel,
                      "note": "same step, total work fixed at the N=1 batch; compare shapes_per_s with the N=1 `value`. "
                              "Limiter at small per-GPU batch: FPS (one CTA per cloud, time independent of B) and the "
                              "launch-latency-bound graph nodes of the deep levels"}
    fast_ddpm = None
    if not args.no_fast_ddpm and args.engine == "fused":
        with torch.no_grad():
            fast_ddpm = fast_ddpm_lines(net, dev, B, dh, cond, label, rank, world, with_cd=(rank == 0 and world == 1))
            net.enable_fused(True, use_tf32=tf32, use_graph=not args.no_graph, fuse_cold=fuse_cold)
    legs = None
    if world == 1 and not args.no_gpu_reference and args.engine == "fused":
        with torch.no_grad():
            legs = baseline_legs(net, dev, B, dh, cond, label, xT_h, steps=5)
    eval_kernels = geometry_kernels = None
    if rank == 0 and not args.no_eval_kernels:
        with torch.no_grad():
            eval_kernels = eval_kernel_lines(dev, clock_info.get("sm_mhz"))
            geometry_kernels = geometry_kernel_lines(dev, clock_info.get("sm_mhz"))
    if rank != 0:
        return
    value = world * B / (T_CHAIN * ms_per_step / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    own_ms = sum(d["ms"] for d in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"]) if agg else (None, None)
    roofline = None
    tf32_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0      # TF32 dense = half the measured bf16 rate (TFLOP/s)
    if top[0]:
        d = top[1]
        per_launch_ms = d["ms"] / d["calls"]
        achieved = (d["bytes"] / d["calls"]) / (per_launch_ms * 1e-3) / 1e9 if d["bytes"] else None
        tflops = (d.get("flops", 0) / (d["ms"] * 1e-3) / 1e12) if d.get("flops") else None
        # both tensor-core kernels of the step against both roofs: the per-layer GEMMs are HBM-bound by design (AI 8-60
        # flop/B); the fused stage kernel reads almost nothing and trades HBM bytes for recomputation, so neither roof is
        # near -- it is bound by the epilogue warps' issue rate, which is why both fractions are given
        both = {}
        for name in ("pdr_gemm_fused", "pdr_stage_chain"):
            if name in agg and agg[name]["ms"] > 0:
                k = agg[name]
                both[name] = {"ms_per_step": round(k["ms"], 4), "launches": k["calls"],
                              "algorithmic_GB": round(k["bytes"] / 1e9, 3), "GFLOP": round(k.get("flops", 0) / 1e9, 1),
                              "hbm_frac": (k["bytes"] / (k["ms"] * 1e-3) / 1e9) / hbm_peak,
                              "tensor_frac": (k.get("flops", 0) / (k["ms"] * 1e-3) / 1e12) / tf32_peak}
        roofline = {"kernel": top[0], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "tflops": tflops, "tensor_peak_tf32_tflops": tf32_peak,
                    "tensor_frac": (tflops / tf32_peak) if tflops else None,
                    "frac": (achieved / hbm_peak) if achieved else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json (burst hbm_gbs; bf16_tflops_sustained / 2 for TF32)" if peaks else "fallback 6.65 TB/s",
                    "launches_per_step": d["calls"], "ms_per_launch": per_launch_ms,
                    "share_of_step": d["ms"] / ms_per_step,
                    "own_kernels_share_of_step": own_ms / ms_per_step,
                    "tensor_core_kernels": both,
                    "per_kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}}
    if roofline:
        # dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), not measured live
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            roofline["traffic"] = tr[roofline["kernel"]]["dram_bytes_per_launch"]
            roofline["traffic_source"] = tr[roofline["kernel"]]["source"]
        except Exception:
            pass
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads(args.cpu_threads)
        dt = cpu_step_seconds(args.cpu_batch, CPU_BASELINE_STEPS, 1)
        cpu_baseline = {"value": args.cpu_batch / (T_CHAIN * dt), "unit": "shapes/s", "cores": cores, "kind": "port",
                        "sample": "%d shape(s) x %d warm steps of the same denoise step on the host "
                                  "(oracle C/OpenMP ops + CPU torch MLPs, %d threads of %d cores); %.2f s/step"
                                  % (args.cpu_batch, CPU_BASELINE_STEPS, cores, os.cpu_count() or 0, dt)}
    line = {
        "metric": "ddpm_shapes_per_sec_T1000", "value": value, "unit": "shapes/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if tf32 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "engine": args.engine, "cuda_graph": (args.engine == "fused" and not args.no_graph),
                   "batch_per_gpu": B, "T": T_CHAIN, "step": "one warm reverse step "
                   "(eps_theta + posterior update, device Philox noise)", "cold_ms": cold_ms,
                   "first_call_ms": first_call_ms, "cold_path": "compiled condition program" if fuse_cold else "torch modules",
                   "l2": "per-step activation working set (>1 GB at B=32) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": "dp%d: shapes sharded by rank, no collective inside the chain, one final all_gather" % world},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clock_info, "e2e": e2e,
        "gpu_launches": launches, "eval_kernels": eval_kernels, "geometry_kernels": geometry_kernels,
        "fast_ddpm": fast_ddpm, "strong_scaling": strong,
        "fp32_simt": (legs or {}).get("fp32_simt"), "gpu_reference": (legs or {}).get("gpu_reference"),
    }
    if legs and legs.get("gpu_reference", {}).get("ms_per_step"):
        line["gpu_reference"]["ours_over_reference_design"] = legs["gpu_reference"]["ms_per_step"] / ms_per_step
    print(json.dumps(line), flush=True, file=JSON_OUT[0])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
