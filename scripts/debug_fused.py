import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from point_diffusion_refinement_b200 import configs
from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
from point_diffusion_refinement_b200.fused import FusedDenoiser
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = C.fill_parameters_(PointNet2CloudCondition(configs.ddpm_pointnet_config()).eval(), seed=1).to(dev)
x, cond, ts, label = [t.to(dev) for t in C.denoiser_inputs(B, 2048, 3072, seed=3)]
with torch.no_grad():
    net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
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
    print("all", len(eng.ops), "ops ok at B =", B, "mem GB", torch.cuda.max_memory_allocated() / 1e9)
