"""ctypes binding of include/nrf_b200.h (the C ABI of libnrf_b200.so).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NRF_LIB_PATH") or os.path.join(HERE, "libnrf_b200.so")   # override: A/B builds

NRF_PREC_BF16 = 0
NRF_PREC_FP32 = 1
NRF_PREC_FP16 = 2
NRF_PREC_BF16X3 = 3
NRF_MAX_BLOCKS = 8

c_f32p = C.c_void_p
c_ptr = C.c_void_p


class NrfGemm(C.Structure):
    _fields_ = [("A", c_ptr * 3), ("K", C.c_int * 3), ("lda", C.c_int * 3),
                ("B", c_ptr), ("ldb", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("n_store", C.c_int),
                ("bias", c_ptr),
                ("mask_src", c_ptr), ("ldmask", C.c_int),
                ("resid", c_ptr), ("ldr", C.c_int),
                ("out_act", c_ptr), ("ldact", C.c_int), ("relu_act", C.c_int),
                ("out_act2", c_ptr), ("ldact2", C.c_int), ("relu_act2", C.c_int),
                ("out_f32", c_ptr), ("ldo", C.c_int)]


class NrfCompositeReuse(C.Structure):
    _fields_ = [("field_new", c_ptr), ("perm", c_ptr), ("n_first", C.c_int), ("d_field_new", c_ptr)]


_PA = c_ptr * NRF_MAX_BLOCKS


class NrfMlpParams(C.Structure):
    _fields_ = [("d_in", C.c_int), ("d_latent", C.c_int), ("d_hidden", C.c_int), ("d_out", C.c_int),
                ("n_blocks", C.c_int), ("n_lin_z", C.c_int),
                ("lin_in_w", c_ptr), ("lin_in_b", c_ptr), ("lin_out_w", c_ptr), ("lin_out_b", c_ptr),
                ("fc0_w", _PA), ("fc0_b", _PA), ("fc1_w", _PA), ("fc1_b", _PA),
                ("lin_z_w", _PA), ("lin_z_b", _PA)]


class NrfMlpGrads(C.Structure):
    _fields_ = [("lin_in_w", c_ptr), ("lin_in_b", c_ptr), ("lin_out_w", c_ptr), ("lin_out_b", c_ptr),
                ("fc0_w", _PA), ("fc0_b", _PA), ("fc1_w", _PA), ("fc1_b", _PA),
                ("lin_z_w", _PA), ("lin_z_b", _PA), ("deterministic", C.c_int), ("d_last", c_ptr),
                ("touch_flags", c_ptr), ("dlatent_ready_event", c_ptr)]


class NrfMlpSizes(C.Structure):
    _fields_ = [("kin_pad", C.c_int), ("dout_pad", C.c_int), ("packed_bytes", C.c_int64),
                ("fwd_bytes_per_sample", C.c_int64), ("bwd_bytes_per_sample", C.c_int64),
                ("bwd_fixed_bytes", C.c_int64)]


_i, _f, _p, _i64 = C.c_int, C.c_float, c_ptr, C.c_int64

_SIGNATURES = {
    "nrf_raygen": [_p, _i, _i, _i, _f, _f, _f, _f, _f, _f, _p, _p, _p],
    "nrf_raygen_ex": [_p, _i, _i, _i, _f, _f, _f, _f, _f, _f, _p, _p, _i, _p],
    "nrf_sample_coarse": [_p, _i, _i, _p, _p, _i, _p, _p],
    "nrf_sample_fine": [_p, _p, _p, _i, _i, _p, _p, _i, _i, _p, _i, _p, _p],
    "nrf_sort_rows": [_p, _i, _i, _p, _p],
    "nrf_volume_to_channels_last": [_p, _p, _i, _i, _i64, _p],
    "nrf_mark_voxels": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "nrf_volume_to_channels_last_marked": [_p, _p, _i, _i, _i64, _p, _i, _p],
    "nrf_volume_to_channels_first": [_p, _p, _i, _i, _i64, _p],
    "nrf_encode_points": [_p, _p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _i, _f, _p, _i, _i, _p, _p],
    "nrf_scatter_volume_grad": [_p, _p, _i, _i, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _p],
    "nrf_composite_fwd": [_p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, C.POINTER(NrfCompositeReuse), _p],
    "nrf_composite_bwd": [_p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p,
                          C.POINTER(NrfCompositeReuse), _i, _p],
    "nrf_gemm": [C.POINTER(NrfGemm), _i, _p],
    "nrf_wgrad": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p],
    "nrf_mlp_sizes": [C.POINTER(NrfMlpParams), _i, C.POINTER(NrfMlpSizes)],
    "nrf_mlp_pack": [C.POINTER(NrfMlpParams), _i, _p, _p],
    "nrf_mlp_fwd": [C.POINTER(NrfMlpParams), _p, _i, _p, _i64, _p, _p, _p],
    "nrf_mlp_fwd_touch": [C.POINTER(NrfMlpParams), _p, _i, _p, _i64, _p, _p, _p, _p],
    "nrf_mlp_fwd_layered": [C.POINTER(NrfMlpParams), _p, _i, _p, _i64, _p, _p, _p],
    "nrf_mlp_fused_supported": [C.POINTER(NrfMlpParams), _i],
    "nrf_mlp_bwd": [C.POINTER(NrfMlpParams), _p, _i, _p, _i64, _p, _p, C.POINTER(NrfMlpGrads), _p, _p, _p],
    "nrf_mlp_bwd_layered": [C.POINTER(NrfMlpParams), _p, _i, _p, _i64, _p, _p, C.POINTER(NrfMlpGrads), _p, _p, _p],
}
_SIGNATURES["nrf_scatter_volume_grad_sorted"] = [_p, _p, _i, _i, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _p]
_SIGNATURES["nrf_render_loss"] = [_p, _p, _p, _p, _i, _i, _i, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p]
_SIGNATURES["nrf_voxelize"] = [_p, _p, _i, _i, _i, _p, _i, _p, _p, _p]
_SIGNATURES["nrf_scatter_volume_grad_merged"] = [_p, _i, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i,
                                                 _p, _p, _p]
_SIGNATURES["nrf_encode_points_touch"] = [_p, _p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _i, _f, _p, _i, _i, _p, _p, _p]
_SIGNATURES["nrf_rows_gather"] = [_p, _i, _i, _i64, _p, _i64, _p, _p]
_SIGNATURES["nrf_rows_update"] = [_p, _i, _i, _i64, _p, _i64, _p, _i, _p]
_SIGNATURES["nrf_rows_merge"] = [_p, _i, _i, _i64, _i, _p, _p, _i64, _p, _i, _i, _p]
_SIGNATURES["nrf_timing_begin"] = []
_SIGNATURES["nrf_timing_end"] = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
EXPORTS = sorted(list(_SIGNATURES) + ["nrf_version", "nrf_last_error", "nrf_wgrad_workspace_bytes",
                                      "nrf_launch_count", "nrf_scatter_sorted_workspace_bytes",
                                      "nrf_voxelize_workspace_bytes"])
TIMING_CATEGORIES = ["gemm_tc", "wgrad_tc", "encode", "composite_fwd", "composite_bwd", "scatter", "transpose",
                     "colsum", "sampling", "simt", "misc", "fused_fwd", "fused_bwd"]

_lib = None


class NrfError(RuntimeError):
    pass


def load():
    """Loads libnrf_b200.so (built by __graft_entry__.build()); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NrfError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.nrf_version.restype = C.c_char_p
    lib.nrf_last_error.restype = C.c_char_p
    lib.nrf_wgrad_workspace_bytes.argtypes = [_i, _i]
    lib.nrf_wgrad_workspace_bytes.restype = C.c_int64
    lib.nrf_scatter_sorted_workspace_bytes.argtypes = [_i64, _i, _i64]
    lib.nrf_scatter_sorted_workspace_bytes.restype = C.c_int64
    lib.nrf_voxelize_workspace_bytes.argtypes = [_i, _i, _i]
    lib.nrf_voxelize_workspace_bytes.restype = C.c_int64
    lib.nrf_launch_count.argtypes = []
    lib.nrf_launch_count.restype = C.c_int64
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nrf_last_error().decode(errors="replace")
        raise NrfError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    """The current CUDA stream of the current device as a void* (what every C-ABI call is enqueued on).  The raw-handle
    query is ~30 x cheaper than building a torch.cuda.Stream object, and this runs once per kernel launch (55 times in
    a step at the reference's training shape, which is host-bound)."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().nrf_launch_count())


def timing_begin():
    check(load().nrf_timing_begin(), "nrf_timing_begin")


def timing_end():
    """-> {category: (milliseconds, launches)} for every kernel launched since timing_begin()."""
    n = len(TIMING_CATEGORIES)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(load().nrf_timing_end(ms, cnt), "nrf_timing_end")
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(TIMING_CATEGORIES)}
