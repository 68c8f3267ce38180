"""EMD_distance timing at BASELINE config 4 (B = 256, 2048 x 2048) and at the evaluation batch (B = 32) -- run once per
PDR_EMD_CLUSTER value (the variable is read once per process)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_diffusion_refinement_b200.emd import EMD_distance
g = torch.Generator().manual_seed(7)
out = {"cluster": os.environ.get("PDR_EMD_CLUSTER", "auto")}
for B in (256, 32):
    a = torch.rand(B, 2048, 3, generator=g).cuda(); b = torch.rand(B, 2048, 3, generator=g).cuda()
    em = EMD_distance()
    for _ in range(2): r = em(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): r = em(a, b)
    e1.record(); torch.cuda.synchronize()
    out["B%d_ms" % B] = e0.elapsed_time(e1) / 3
    out["B%d_checksum" % B] = float(r.double().sum())
print(json.dumps(out))
