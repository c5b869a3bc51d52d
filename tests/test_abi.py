"""CPU suite: the C-ABI library builds, loads and exports every symbol include/rsm.h declares,
and the ctypes prototypes in realtime_stereo_matcher_b200/_lib.py cover exactly those symbols.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rsm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from realtime_stereo_matcher_b200 import build
    return ctypes.CDLL(build.build())


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rsm.h but not exported"


def test_ctypes_prototypes_cover_header():
    from realtime_stereo_matcher_b200 import _lib
    assert sorted(list(_lib.SIGNATURES) + list(_lib.OTHER_SYMBOLS)) == declared_symbols()


def test_version_and_error_strings(lib):
    lib.rsm_version.restype = ctypes.c_int
    lib.rsm_last_error.restype = ctypes.c_char_p
    from realtime_stereo_matcher_b200 import _lib
    assert lib.rsm_version() == _lib.RSM_VERSION
    assert lib.rsm_last_error(0) == b"ok"
    assert b"C % G" in lib.rsm_last_error(1)
    assert b"unknown" in lib.rsm_last_error(999)


def test_argument_validation_without_gpu():
    """Shape / dtype / null checks run before any CUDA call, so they are testable on CPU."""
    from realtime_stereo_matcher_b200 import _lib as L
    lib = L.load()
    f = L.RsmFeat(None, 0, 0, 0, 0)
    assert lib.rsm_concat_fwd(f, f, None, -1, 1, 1, 1, 1, 0, 0, None) == 1        # invalid shape
    assert lib.rsm_concat_fwd(f, f, None, 0, 4, 4, 4, 4, 0, 0, None) == 0         # empty batch: no-op
    assert lib.rsm_concat_fwd(f, f, None, 1, 1, 1, 1, 1, 0, 0, None) == 3         # null pointer
    assert lib.rsm_groupwise_fwd(f, f, None, 1, 6, 1, 1, 1, 4, 0, 0, 0, None) == 1  # C % G != 0
    assert lib.rsm_regress_fwd(None, 1, 0, 1, 1, 0, L.RsmRegressOut(), 0, None) == 1  # empty D
    assert lib.rsm_upsample_regress_bwd_workspace(2, 48, 8, 8) == 2 * 48 * 8 * 8 * 4


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    import realtime_stereo_matcher_b200 as rsm
    x = torch.zeros(1, 4, 2, 8)
    for fn in (lambda: rsm.concat_volume(x, x, 2), lambda: rsm.interweave(x, x),
               lambda: rsm.inner_product_volume(x, x, 2), lambda: rsm.groupwise_volume(x, x, 2, 2),
               lambda: rsm.difference_volume(x, x, 2), lambda: rsm.soft_argmax(x),
               lambda: rsm.upsample_regress(x, 8, 4, 16), lambda: rsm.inner_product_regress(x, x, 2)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "realtime_stereo_matcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
