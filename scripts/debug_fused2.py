import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from point_diffusion_refinement_b200 import util
from point_diffusion_refinement_b200.fused import FusedDenoiser
dev = torch.device("cuda", 0)
B = 32
net = bench.make_net(dev)
cond_h, label_h, xT_h = bench.make_inputs(B, seed=100)
cond, label, x = cond_h.to(dev), label_h.to(dev), xT_h.to(dev)
ts = torch.full((B,), 999.0, device=dev)
rng = util.DeviceNoise(seed=1234)
with torch.no_grad():
    eps = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
    print("cold eps finite", torch.isfinite(eps).all().item(), eps.abs().max().item())
    rng.affine_update(x, eps.contiguous(), 1.01, -0.01, 0.1)
    print("x finite", torch.isfinite(x).all().item(), x.abs().max().item())
    eng = FusedDenoiser(net, B, 2048, use_graph=False)
    eng.set_condition(net._cond_state, label)
    eng.x_in.copy_(x); eng.ts_in.copy_(ts)
    torch.cuda.synchronize()
    for i, (op, meta) in enumerate(zip(eng.ops, eng.meta)):
        op()
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("FAILED at op", i, meta, str(e)[:100]); sys.exit(1)
    print("all ops ok; eps finite", torch.isfinite(eng.eps_out).all().item())
