"""Runs the compiled warm step (B = 32) eagerly and brackets the ops [LO, HI) of the program with cudaProfilerStart / Stop, for
`ncu --profile-from-start off`.  Usage: python scripts/prof_engine_ops.py LO HI   (op indices as in bench.py --dump-ops)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from point_diffusion_refinement_b200 import configs  # noqa: E402
from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition  # noqa: E402
from tests import common as C  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
x, cond, ts, label = [t.to(dev) for t in C.denoiser_inputs(32, 2048, 3072, seed=3)]
net = C.fill_parameters_(PointNet2CloudCondition(configs.ddpm_pointnet_config()).eval(), seed=1).to(dev)
net.enable_fused(True, use_tf32=True, use_graph=False, fuse_cold=True)
with torch.no_grad():
    cold = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
    x2 = x + 0.05 * cold
    for _ in range(2):
        net(x2, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
    eng = net._fused_engine
    torch.cuda.synchronize()
    names = [n for n, _ in eng.meta]
    print("ops", len(eng.ops), "profiling", [(k, names[k], eng.meta[k][1].get("M"), eng.meta[k][1].get("N"), eng.meta[k][1].get("K"))
                                             for k in range(lo, hi)])
    for k, op in enumerate(eng.ops):
        if k == lo:
            torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
        op()
        if k == hi - 1:
            torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
    torch.cuda.synchronize()
print("done")
