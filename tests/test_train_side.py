"""Training-side pieces of SURVEY 8f rank 4 (next rows): the gradient all-reduce over a world-2 gloo group on CPU, and (GPU)
the Chamfer backward and three_interpolate backward kernels against autograd."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from point_diffusion_refinement_b200 import dist as pd
    pd.init_from_env(backend="gloo")
    torch.manual_seed(100 + rank)                                   # ranks start from different weights ...
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    pd.apply_gradient_allreduce(net)                                # ... and are broadcast from rank 0
    w0 = net[0].weight.detach().clone()
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()                                         # hook: gradients averaged at the end of backward
    q.put((rank, w0, net[0].weight.grad.clone(), net[2].bias.grad.clone()))
    torch.distributed.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """reference pointnet2/distributed.py:94-146: parameters broadcast from rank 0, gradients averaged over the ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, g_a, b_a), (_, w_b, g_b, b_b) = res
    assert torch.equal(w_a, w_b)                                    # broadcast
    assert torch.equal(g_a, g_b) and torch.equal(b_a, b_b)          # same averaged gradient on both ranks
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    grads = []
    for r in range(world):
        ref.zero_grad()
        ref(torch.full((4, 5), float(r + 1))).sum().backward()
        grads.append(ref[0].weight.grad.clone())
    torch.testing.assert_close(g_a, (grads[0] + grads[1]) / 2, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_chamfer3d_backward_matches_autograd():
    """chamfer_3DFunction.backward (chamfer3D.cu:155-196) against autograd through the same argmin assignment."""
    from point_diffusion_refinement_b200.dist_chamfer_3D import chamfer_3DDist
    g = torch.Generator().manual_seed(0)
    for b, n, m in ((2, 300, 500), (3, 2048, 1024), (1, 5, 3000)):
        a = torch.rand(b, n, 3, generator=g).cuda().requires_grad_(True)
        c = torch.rand(b, m, 3, generator=g).cuda().requires_grad_(True)
        d1, d2, i1, i2 = chamfer_3DDist()(a, c)
        w1, w2 = torch.rand(b, n, generator=g).cuda(), torch.rand(b, m, generator=g).cuda()
        ((d1 * w1).sum() + (d2 * w2).sum()).backward()
        a2, c2 = a.detach().clone().requires_grad_(True), c.detach().clone().requires_grad_(True)
        e1 = ((a2 - torch.gather(c2, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))) ** 2).sum(-1)
        e2 = ((c2 - torch.gather(a2, 1, i2.long().unsqueeze(-1).expand(-1, -1, 3))) ** 2).sum(-1)
        torch.testing.assert_close(d1, e1, rtol=1e-5, atol=1e-7)
        ((e1 * w1).sum() + (e2 * w2).sum()).backward()
        torch.testing.assert_close(a.grad, a2.grad, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(c.grad, c2.grad, rtol=1e-4, atol=1e-6)
        g1 = a.grad.clone()
        a.grad = None; c.grad = None
        d1, d2, _, _ = chamfer_3DDist()(a, c)
        ((d1 * w1).sum() + (d2 * w2).sum()).backward()
        assert torch.equal(a.grad, g1)                              # deterministic (no atomics)


@pytest.mark.gpu
def test_three_interpolate_backward_matches_autograd():
    """three_interpolate_grad (interpolate_gpu.cu:104-154) through the autograd Function of pointnet2_utils."""
    from point_diffusion_refinement_b200 import pointnet2_utils as U
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(2, 9, 40, generator=g).cuda().requires_grad_(True)
    idx = torch.randint(0, 40, (2, 100, 3), generator=g).int().cuda()
    w = torch.rand(2, 100, 3, generator=g)
    w = (w / w.sum(2, keepdim=True)).cuda()
    out = U.three_interpolate(feats, idx, w)
    up = torch.randn(out.shape, generator=g).cuda()
    (out * up).sum().backward()
    f2 = feats.detach().clone().requires_grad_(True)
    gathered = torch.gather(f2.unsqueeze(2).expand(-1, -1, 100, -1), 3, idx.long().unsqueeze(1).expand(-1, 9, -1, -1))
    ref = (gathered * w.unsqueeze(1)).sum(-1)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    (ref * up).sum().backward()
    torch.testing.assert_close(feats.grad, f2.grad, rtol=1e-4, atol=1e-5)
