"""HBM ceilings by access mix on this GPU (CUDA events, best of 10): write-only (memset), read-only (reduction),
copy (1:1) and a 1:3 read:write mix -- the mixes the GEMMs of the step actually have.  Prints one JSON line."""
import json
import torch

dev = torch.device("cuda")
n = 1 << 30                                   # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
a.fill_(1.0); b.fill_(2.0)


def best(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return nbytes / t / 1e6


q = n // 4
h = n // 2
out = {
    "write_only_zero_GBps": best(lambda: a.zero_(), 4 * n),
    "write_only_const_GBps": best(lambda: a.fill_(1.2345), 4 * n),
    "write_only_arange_GBps": best(lambda: torch.arange(0, n, dtype=torch.float32, out=a), 4 * n),
    # read n/2 once, write it to two destinations: 1 read : 2 writes
    "read1_write2_GBps": best(lambda: (b[:h].copy_(a[:h]), b[h:].copy_(a[:h])), 4 * 3 * h),
    "read_only_GBps": best(lambda: torch.sum(a), 4 * n),
    "copy_GBps": best(lambda: b.copy_(a), 8 * n),
    # read n/4, write 3n/4 ... expand of a quarter-sized source into three quarter-sized destinations
    "read1_write3_GBps": best(lambda: b[:3 * q].view(3, q).copy_(a[:q].unsqueeze(0).expand(3, q)), 4 * 4 * q),
    "read2_write1_GBps": best(lambda: torch.add(a[:q], a[q:2 * q], out=b[:q]), 4 * 3 * q),
}
# incompressible data through our own store paths (pdr_probe_hbm): fill_/zero_ write a constant
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_diffusion_refinement_b200 import _lib
L = _lib.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def probe(mode):
    rc = L.pdr_probe_hbm(mode, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_size_t(n), st)
    assert rc == 0, L.pdr_last_error_string()


out["write_only_hash_stg128_GBps"] = best(lambda: probe(0), 4 * n)
out["read1_write3_hash_stg128_GBps"] = best(lambda: probe(1), 4 * n * 4 // 3)
out["write_only_hash_tma_bulk_GBps"] = best(lambda: probe(2), 4 * n)
chk = a[:4096].clone()
assert torch.isfinite(chk).all() and chk.min() >= 1 and chk.max() < 2 and chk.unique().numel() > 4000
print(json.dumps(out))
