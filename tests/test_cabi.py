"""CPU: the C-ABI library loads and exports every symbol include/pdr_b200.h declares (no compute)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cuda_lib):
    from point_diffusion_refinement_b200 import _lib
    declared = _declared()
    assert len(declared) >= 25
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (pdr_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_lib.SIGNATURES), (sorted(set(declared) ^ set(_lib.SIGNATURES)))
    for name in declared:
        assert getattr(cuda_lib, name) is not None


def test_identity_and_error_plumbing(cuda_lib):
    assert cuda_lib.pdr_version() == 100 and cuda_lib.pdr_built_for_sm() == 100
    assert cuda_lib.pdr_fps_max_onchip_points() == 16384
    assert cuda_lib.pdr_emd_workspace_bytes(2, 10, 20) >= 2 * 2 * 30 * 4
    # invalid sizes are rejected before anything touches the device
    rc = cuda_lib.pdr_ball_query(1, 0, 1, 0.1, 4, None, None, None, None, None)
    assert rc == -1 and b"ball_query" in cuda_lib.pdr_last_error_string()
    rc = cuda_lib.pdr_knn_points(1, 4, 4, 100, None, None, None, None, None)
    assert rc == -1 and b"K=100" in cuda_lib.pdr_last_error_string()


def test_sass_is_sm100(cuda_lib):
    from point_diffusion_refinement_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: CPU tensors are rejected like the reference's AT_ASSERT(false, "CPU not supported")."""
    import pytest
    import torch
    from point_diffusion_refinement_b200 import _ext, knn, emd_cuda
    x = torch.rand(1, 8, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        _ext.furthest_point_sampling(x, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        _ext.ball_query(x, x, 0.1, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        knn.knn_points(x, x, K=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        emd_cuda.emd_cost_forward(x, x)


def test_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "point_diffusion_refinement_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src and "cpu_oracle" not in src, fn


def test_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """PdrGemmArgs / PdrGnArgs are passed by address: the ctypes mirrors in fused.py must have the header's layout
    field for field (gcc's offsetof/sizeof against ctypes')."""
    import ctypes
    from point_diffusion_refinement_b200.fused import GemmArgs, GnArgs, GnSource
    header = os.path.join(ROOT, "include", "pdr_b200.h")
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)

    def c_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names.append(re.sub(r"\[.*\]", "", decl.split()[-1].lstrip("*")))
        return names

    src = ['#include <stddef.h>', '#include <stdio.h>', '#include "pdr_b200.h"', 'int main(void) {']
    want = {}
    for cname, mirror in (("PdrGemmArgs", GemmArgs), ("PdrGnSource", GnSource), ("PdrGnArgs", GnArgs)):
        fields = c_fields(cname)
        assert fields == [f[0] for f in mirror._fields_], (cname, fields, [f[0] for f in mirror._fields_])
        src.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        want[cname] = ctypes.sizeof(mirror)
        for f in fields:
            src.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
            want["%s.%s" % (cname, f)] = getattr(mirror, f).offset
    src += ["  return 0;", "}"]
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(cfile)])
    got = dict((k, int(v)) for k, v in (l.split() for l in subprocess.check_output([exe], text=True).splitlines()))
    assert got == want, sorted(k for k in want if got.get(k) != want[k])
