"""CPU: the N>1 path (shard by shapes, no collective during sampling, one final all_gather) with
world_size 2 over gloo."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from point_diffusion_refinement_b200 import dist as pd
    r, w, _ = pd.init_from_env(backend="gloo")

    def sample(start, stop):  # stand-in sampler: shape id in every coordinate
        return torch.arange(start, stop, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3).contiguous()

    out = pd.sample_sharded(sample, total, r, w)
    q.put((rank, out[:, 0, 0].tolist()))
    torch.distributed.destroy_process_group()


def test_shard_range_covers_everything():
    from point_diffusion_refinement_b200.dist import shard_range
    for total in (0, 1, 7, 32, 33, 41600):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_uneven():
    world, total = 2, 7   # ranks own 4 and 3 shapes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert results[r] == [float(i) for i in range(total)]
