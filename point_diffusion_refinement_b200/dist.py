"""Multi-GPU plumbing for the sampling path: one process per GPU, shapes sharded across ranks, no
collective during the T reverse steps, ONE all_gather of the generated clouds (and of the per-shape
metrics) at the end over NCCL/NVLink.

Replaces the reference's Popen fan-out + filesystem gather (pointnet2/generate_samples_distributed.py:
26-97,188-201) and the per-step nn.DataParallel scatter/gather (pointnet2/completion_eval.py:113-118).
Sharding follows the reference's contiguous-range rule (mvp_dataloader/mvp_dataset.py:152-198).
"""
import os

import torch
import torch.distributed as dist


def rank():
    """Rank in the default process group, 0 without one."""
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def shard_range(total, rank, world_size):
    """Contiguous [start, stop) of `total` shapes owned by `rank`; the first `total % world` ranks get one
    extra shape."""
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT (torchrun).  Returns
    (rank, world_size, local_rank).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)      # binds the communicator to this rank's GPU
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def all_gather_shapes(local, counts=None):
    """Concatenate per-rank tensors (n_r, ...) along dim 0 on every rank.  Ranks may hold different n_r
    (pass `counts`, the list of n_r, or let it be exchanged first)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        counts = [int(v.item()) for v in ns]
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def sample_sharded(sample_fn, total_shapes, rank, world_size):
    """Run `sample_fn(start, stop) -> (n_local, N, 3)` on this rank's shard and gather everything."""
    start, stop = shard_range(total_shapes, rank, world_size)
    local = sample_fn(start, stop)
    counts = [shard_range(total_shapes, r, world_size) for r in range(world_size)]
    return all_gather_shapes(local, counts=[b - a for a, b in counts])


# ---- training-side collectives (reference pointnet2/distributed.py:67-146) -------------------------
def broadcast_parameters(module, src=0):
    """Every tensor of the state_dict from `src` to all ranks (distributed.py:104-107)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return module
    for t in module.state_dict().values():
        if torch.is_tensor(t):
            dist.broadcast(t, src)
    return module


def all_reduce_gradients(module):
    """Average the gradients over the ranks: one flat bucket per dtype, one all_reduce per bucket over NCCL / NVLink
    (distributed.py:109-133 without the per-backward callback plumbing).  Returns the number of bytes reduced."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    buckets = {}
    for p in module.parameters():
        if p.requires_grad and p.grad is not None:
            buckets.setdefault(p.grad.dtype, []).append(p.grad.data)
    total = 0
    for grads in buckets.values():
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= world
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        total += flat.numel() * flat.element_size()
    return total


def apply_gradient_allreduce(module):
    """The reference's wrapper (distributed.py:94-146): broadcast the parameters once, then average the gradients at
    the end of every backward pass that follows a forward pass."""
    broadcast_parameters(module)
    module.needs_reduction = False

    def reduce_once():
        if module.needs_reduction:
            module.needs_reduction = False
            all_reduce_gradients(module)

    def hook(*_):
        torch.autograd.Variable._execution_engine.queue_callback(reduce_once)

    for p in module.parameters():
        if p.requires_grad:
            p.register_hook(hook)
    module.register_forward_hook(lambda m, i, o: setattr(m, "needs_reduction", True))
    return module
