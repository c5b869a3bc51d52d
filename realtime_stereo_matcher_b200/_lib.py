"""ctypes binding of librsm_b200.so (the C ABI declared in include/rsm.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module on a
machine without the built library raises; calling an op with a non-CUDA tensor raises.
ctypes releases the GIL for the duration of each foreign call, so nn.DataParallel worker
threads (reference train_stereo.py:139) enqueue on their own devices concurrently.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librsm_b200.so")

RSM_F32, RSM_F16, RSM_BF16 = 0, 1, 2
RSM_REDUCE_SUM, RSM_REDUCE_MEAN = 0, 1
_DTYPES = {torch.float32: RSM_F32, torch.float16: RSM_F16, torch.bfloat16: RSM_BF16}


class RsmFeat(C.Structure):
    _fields_ = [("data", C.c_void_p), ("stride_n", C.c_int64), ("stride_c", C.c_int64),
                ("stride_h", C.c_int64), ("stride_w", C.c_int64)]


class RsmRegressOut(C.Structure):
    _fields_ = [("soft", C.c_void_p), ("argmin", C.c_void_p), ("argmax", C.c_void_p), ("lse", C.c_void_p),
                ("expect", C.c_void_p)]


class RsmV4Weights(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("t1", C.c_void_p), ("w2", C.c_void_p), ("t2", C.c_void_p), ("w3", C.c_void_p),
                ("t3", C.c_void_p), ("w11", C.c_void_p), ("t11", C.c_void_p)]


i64, vp, ci, cf = C.c_int64, C.c_void_p, C.c_int, C.c_float

RSM_REDUCE_WS_DOUBLES = 1184 * 8   # include/rsm.h
RSM_VERSION = 107                  # include/rsm.h; the argument layouts below were written for this ABI

# name -> argtypes, exactly the prototypes of include/rsm.h (tests/test_abi.py checks the header)
SIGNATURES = {
    "rsm_concat_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_concat_bwd": [vp, vp, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_interweave_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_interweave_bwd": [vp, vp, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_inner_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, i64, ci, ci, ci, ci, vp],
    "rsm_inner_bwd": [vp, RsmFeat, RsmFeat, vp, vp, i64, i64, i64, i64, i64, ci, ci, ci, ci, vp],
    "rsm_inner_bwd_profile": [vp, RsmFeat, RsmFeat, vp, vp, i64, i64, i64, i64, i64, ci, ci, ci, ci, vp, vp],
    "rsm_groupwise_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp],
    "rsm_groupwise_bwd": [vp, RsmFeat, RsmFeat, vp, vp, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp],
    "rsm_difference_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, i64, cf, ci, ci, vp],
    "rsm_difference_bwd": [vp, vp, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_shift_interweave_fwd": [RsmFeat, RsmFeat, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_shift_interweave_bwd": [vp, vp, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_v4_volume_fwd": [RsmFeat, RsmFeat, RsmV4Weights, vp, vp, i64, i64, i64, i64, i64, ci, ci, ci, vp],
    "rsm_v4_volume_fwd_profile": [RsmFeat, RsmFeat, RsmV4Weights, vp, vp, i64, i64, i64, i64, i64, ci, ci, ci, vp, vp],
    "rsm_warp_fwd": [vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, vp],
    "rsm_warp_bwd": [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, vp],
    "rsm_pfm_write": [C.c_char_p, vp, i64, i64, ci, C.c_double, ci],
    "rsm_pfm_read_header": [C.c_char_p, vp, vp, vp, vp, vp],
    "rsm_pfm_read": [C.c_char_p, vp, i64, i64, ci, ci],
    "rsm_seqloss_fwd": [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, cf, ci, ci, ci, vp],
    "rsm_seqloss_bwd": [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, cf, ci, ci, ci, vp],
    "rsm_flow_metrics": [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_prepare_fwd": [vp, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_prepare_bwd": [vp, vp, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_finalize_fwd": [vp, vp, i64, i64, i64, i64, i64, i64, i64, cf, ci, ci, ci, vp],
    "rsm_finalize_bwd": [vp, vp, i64, i64, i64, i64, i64, i64, i64, cf, ci, ci, ci, vp],
    "rsm_regress_fwd": [vp, i64, i64, i64, i64, ci, RsmRegressOut, ci, vp],
    "rsm_regress_bwd": [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_expect_fwd": [vp, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_expect_bwd": [vp, vp, i64, i64, i64, i64, ci, ci, vp],
    "rsm_upsample_regress_fwd": [vp, i64, i64, i64, i64, i64, i64, i64, ci, RsmRegressOut, ci, vp],
    "rsm_upsample_regress_bwd": [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, ci, vp],
    "rsm_inner_regress_fwd": [RsmFeat, RsmFeat, i64, i64, i64, i64, i64, ci, ci, RsmRegressOut, ci, vp],
    "rsm_inner_regress_fwd_profile": [RsmFeat, RsmFeat, i64, i64, i64, i64, i64, ci, ci, RsmRegressOut, ci, vp, vp],
}
OTHER_SYMBOLS = ("rsm_version", "rsm_last_error", "rsm_upsample_regress_bwd_workspace", "rsm_v4_volume_workspace")

_lib = None


def load() -> C.CDLL:
    """Load librsm_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m realtime_stereo_matcher_b200.build` "
            "(nvcc, sm_100a). realtime_stereo_matcher_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.rsm_version.argtypes = []
    lib.rsm_version.restype = C.c_int
    lib.rsm_last_error.argtypes = [C.c_int]
    lib.rsm_last_error.restype = C.c_char_p
    lib.rsm_upsample_regress_bwd_workspace.argtypes = [i64, i64, i64, i64]
    lib.rsm_upsample_regress_bwd_workspace.restype = i64
    lib.rsm_v4_volume_workspace.argtypes = [i64, i64, i64, i64]
    lib.rsm_v4_volume_workspace.restype = i64
    got = lib.rsm_version()
    if got != RSM_VERSION:
        # the .so is git-ignored and copied between boxes: a stale binary would be called with the wrong layouts
        raise RuntimeError(
            f"{LIB_PATH} reports ABI version {got}, this package was written for {RSM_VERSION}: rebuild it with "
            "`python -m realtime_stereo_matcher_b200.build --force`")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().rsm_last_error(code).decode()
        raise RuntimeError(f"{what} failed with rsm status {code}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"realtime_stereo_matcher_b200 supports float32/float16/bfloat16, got {t.dtype}") from None


def require_cuda(*tensors: torch.Tensor) -> int:
    """All tensors must live on one CUDA device; returns its index."""
    dev = None
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "realtime_stereo_matcher_b200 runs on CUDA (sm_100a) only and has no CPU fallback; "
                f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device.index
        elif t.device.index != dev:
            raise RuntimeError(f"tensors on different devices: cuda:{dev} and {t.device}")
    return dev


def stream_ptr(device: int) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def feat(t: torch.Tensor) -> RsmFeat:
    s = t.stride()
    return RsmFeat(t.data_ptr(), s[0], s[1], s[2], s[3])


def ptr(t):
    return None if t is None else t.data_ptr()
