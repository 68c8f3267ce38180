"""Chamfer_F1 timing at BASELINE config 4 (B = 256, 2048 x 2048 and 16384 x 16384) with checksums."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1
g = torch.Generator().manual_seed(7)
out = {}
cf = Chamfer_F1()
for n in (2048, 16384):
    a = torch.rand(256, n, 3, generator=g).cuda(); b = torch.rand(256, n, 3, generator=g).cuda()
    for _ in range(2): r = cf(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if n == 2048 else 2
    e0.record()
    for _ in range(reps): r = cf(a, b)
    e1.record(); torch.cuda.synchronize()
    out["n%d_ms" % n] = e0.elapsed_time(e1) / reps
    out["n%d_checksum" % n] = [float(t.double().sum()) for t in r]
print(json.dumps(out))
