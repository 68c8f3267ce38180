"""ctypes binding of libpdr_b200.so (the C ABI declared in include/pdr_b200.h).

There is no CPU fallback and no alternative backend: if the library is missing or a call fails the
caller gets an exception.  Nothing in this package imports ``oracle/``.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpdr_b200.so")

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t
_c_u64 = ctypes.c_uint64
_ptr = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES).  Must stay in step with pdr_b200.h;
# tests/test_cabi.py checks every declaration in the header against this table and the .so exports.
SIGNATURES = {
    "pdr_version": [],
    "pdr_last_error_string": [],
    "pdr_built_for_sm": [],
    "pdr_probe_hbm": [_c_int, _ptr, _ptr, _c_size_t, _ptr],
    "pdr_fps_max_onchip_points": [],
    "pdr_furthest_point_sampling": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "pdr_gather_points": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "pdr_gather_points_grad": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "pdr_ball_query": [_c_int, _c_int, _c_int, _c_float, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_group_points": [_c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "pdr_group_points_grad": [_c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "pdr_three_nn": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_three_interpolate": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_three_interpolate_grad": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_knn_points": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_chamfer_f1_workspace_bytes": [_c_int, _c_int, _c_int],
    "pdr_chamfer_f1": [_c_int, _c_int, _c_int, _ptr, _ptr, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                       _c_size_t, _ptr],
    "pdr_nm_distance": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_nm_distance_grad": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_emd_workspace_bytes": [_c_int, _c_int, _c_int],
    "pdr_emd_approxmatch": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_size_t, _ptr],
    "pdr_emd_matchcost": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_size_t, _ptr],
    "pdr_emd_cost": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_size_t, _ptr],
    "pdr_emd_matchcost_backward": [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pdr_ddpm_update": [_c_size_t, _ptr, _ptr, _c_float, _c_float, _c_float, _ptr, _c_u64, _c_u64, _ptr],
    "pdr_affine_noise_update": [_c_size_t, _ptr, _ptr, _c_float, _c_float, _c_float, _ptr, _c_u64, _c_u64,
                                _ptr],
    "pdr_normal_fill": [_c_size_t, _ptr, _c_u64, _c_u64, _ptr],
    "pdr_point_upsample": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _c_float, _c_float, _ptr, _ptr, _ptr],
    # fused denoiser primitives (struct arguments are passed by address)
    "pdr_gemm_tile_rows": [],
    "pdr_gemm_fused": [_ptr, _ptr],
    "pdr_gn_finalize": [_ptr, _ptr],
    "pdr_gn_finalize_batch": [_ptr, _c_int, _ptr],
    "pdr_affine_rows": [_c_int, _c_int, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _c_int, _ptr, _c_int, _ptr, _c_int,
                        _ptr, _c_int, _c_int, _ptr],
    "pdr_attention_pool": [_c_int, _c_int, _c_int, _c_int, _ptr, _c_int, _ptr, _c_int, _ptr, _ptr, _c_int, _ptr, _ptr,
                           _c_int, _c_int, _ptr],
    "pdr_group_ball": [_c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr,
                       _c_int, _ptr],
    "pdr_group_knn": [_c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr],
    "pdr_group_geo_ball": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _c_int, _ptr],
    "pdr_group_geo_knn": [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr],
    "pdr_group_src_rows": [_c_int, _c_int, _c_int, _c_int, _ptr, _c_int, _ptr, _c_int, _ptr, _ptr],
    "pdr_gather_rows": [_c_int, _c_int, _c_int, _c_int, _ptr, _c_int, _ptr, _ptr, _c_int, _c_int, _ptr],
    "pdr_stage_chain_tile_rows": [],
    "pdr_stage_chain": [_ptr, _ptr],
}
_RESTYPES = {
    "pdr_last_error_string": ctypes.c_char_p,
    "pdr_chamfer_f1_workspace_bytes": _c_size_t,
    "pdr_emd_workspace_bytes": _c_size_t,
}

_lib = None
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports kernel launches from this)


class PdrError(RuntimeError):
    pass


def lib():
    """Load libpdr_b200.so or raise.  Never falls back to anything."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdrError(
                "libpdr_b200.so is not built (%s). Run `python -m point_diffusion_refinement_b200.build` "
                "or __graft_entry__.build(); there is no CPU/PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _c_int)
        _lib = handle
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point; raise PdrError with the library's message on failure."""
    global launch_count
    rc = getattr(lib(), name)(*args)
    launch_count += 1
    if rc != 0:
        msg = lib().pdr_last_error_string()
        raise PdrError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))


def stream_ptr(t):
    """cudaStream_t of torch's current stream on t's device (the reference launches on
    at::cuda::getCurrentCUDAStream(), e.g. sampling_gpu.cu:25-26)."""
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def dptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def check_cuda_f32(t, name):
    """CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_CUDA of the reference bindings (utils.h:5-25)."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def check_cuda_i32(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if t.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
