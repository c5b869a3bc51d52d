"""GPU parity of the tcgen05 inner-product adjoint (csrc/rsm_corr_bwd_tc.cu: 16-bit tensors) against the fp32
oracle on the same rounded inputs: both gradients and one gradient alone, channel sum and mean, 16 .. 192 channels
(one and several 64-channel passes), D = 1 .. 64 in one launch and D = 65 .. 200 as chunks of 64 added with TMA
reduce-add, rows narrower than a tile (and narrower than D) and ragged last tiles, the x < d fill
region with non-finite upstream gradients, strided feature views, and the SIMT fall-back for shapes it does not cover.
Reference: autograd through TorchInnerProductCost.forward (cost_volume/inner_product.py:29-41) / make_correlation_volume
(model/mobile_disp_net_c.py:188-205); SURVEY.md 8a "Backward contracts"."""
import numpy as np
import pytest
import torch

import oracle
from golden_io import round_to
from tolerances import RTOL_16

pytestmark = pytest.mark.gpu

DT = {"fp16": torch.float16, "bf16": torch.bfloat16}
SHAPES = [(1, 16, 2, 128, 16), (1, 64, 3, 240, 48), (2, 64, 5, 240, 48), (1, 32, 4, 312, 48), (1, 16, 2, 72, 19), (1, 128, 2, 480, 64),
          (1, 48, 2, 136, 1), (1, 16, 2, 8, 24), (3, 32, 7, 96, 64), (1, 16, 1, 520, 33), (2, 192, 2, 264, 40),
          (1, 64, 2, 480, 96), (1, 128, 2, 480, 192), (2, 32, 3, 200, 65), (1, 16, 2, 72, 130), (1, 16, 2, 312, 200)]


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()
    return m


def _case(shape, dn, seed=5):
    n, c, h, w, d = shape
    rng = np.random.default_rng(seed)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    go = round_to(rng.standard_normal((n, d, h, w)).astype(np.float32), dn)
    return l, r, go


def _check(got, want, dn, d, c, mean):
    atol = RTOL_16[dn] * np.sqrt(d) * 4 / (c if mean else 1)
    np.testing.assert_allclose(got.float().cpu().numpy(), want, atol=atol, rtol=RTOL_16[dn])


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dn", ["bf16", "fp16"])
@pytest.mark.parametrize("mean", [False, True])
def test_inner_bwd_tc(rsm, shape, dn, mean):
    n, c, h, w, d = shape
    l, r, go = _case(shape, dn)
    gl, gr = oracle.inner_product_volume_bwd(go, l, r, mean=mean)
    lt = torch.from_numpy(l).cuda().to(DT[dn]).requires_grad_(True)
    rt = torch.from_numpy(r).cuda().to(DT[dn]).requires_grad_(True)
    gt = torch.from_numpy(go).cuda().to(DT[dn])
    rsm.inner_product_volume(lt, rt, d, mean=mean).backward(gt)
    _check(lt.grad, gl, dn, d, c, mean)
    _check(rt.grad, gr, dn, d, c, mean)
    # one gradient only (the other side is skipped inside the kernel)
    l2 = torch.from_numpy(l).cuda().to(DT[dn]).requires_grad_(True)
    rsm.inner_product_volume(l2, rt.detach(), d, mean=mean).backward(gt)
    _check(l2.grad, gl, dn, d, c, mean)
    r2 = torch.from_numpy(r).cuda().to(DT[dn]).requires_grad_(True)
    rsm.inner_product_volume(lt.detach(), r2, d, mean=mean).backward(gt)
    _check(r2.grad, gr, dn, d, c, mean)


def test_inner_bwd_tc_fill_region_is_ignored(rsm):
    """Upstream gradient entries with x < d belong to the volume's fill region: the reference's slice assignment never
    reads them, so NaN / inf there must not reach either gradient."""
    _fill_region_case(rsm, (2, 32, 3, 200, 48))


def test_inner_bwd_tc_fill_region_is_ignored_chunked(rsm):
    """the same with three disparity chunks: the fill region of a later chunk reaches into the second tile of a row"""
    _fill_region_case(rsm, (1, 32, 2, 328, 160))


def _fill_region_case(rsm, shape):
    n, c, h, w, d = shape
    l, r, go = _case(shape, "bf16", seed=11)
    for dd in range(d):
        go[:, dd, :, :dd] = np.nan if dd % 2 else np.inf
    clean = go.copy()
    for dd in range(d):
        clean[:, dd, :, :dd] = 0.0
    gl, gr = oracle.inner_product_volume_bwd(clean, l, r)
    lt = torch.from_numpy(l).cuda().bfloat16().requires_grad_(True)
    rt = torch.from_numpy(r).cuda().bfloat16().requires_grad_(True)
    rsm.inner_product_volume(lt, rt, d).backward(torch.from_numpy(go).cuda().bfloat16())
    assert torch.isfinite(lt.grad).all() and torch.isfinite(rt.grad).all()
    _check(lt.grad, gl, "bf16", d, c, False)
    _check(rt.grad, gr, "bf16", d, c, False)


def test_inner_bwd_tc_views_and_fallback(rsm):
    n, c, h, w, d = 2, 32, 3, 248, 48
    rng = np.random.default_rng(9)
    lf = round_to(rng.standard_normal((n, c + 16, h, w + 8)).astype(np.float32), "bf16")
    rf = round_to(rng.standard_normal((n, c + 16, h, w + 8)).astype(np.float32), "bf16")
    go = round_to(rng.standard_normal((n, d, h, w)).astype(np.float32), "bf16")
    lt = torch.from_numpy(lf).cuda().bfloat16().requires_grad_(True)
    rt = torch.from_numpy(rf).cuda().bfloat16().requires_grad_(True)
    rsm.inner_product_volume(lt[:, 8:8 + c, :, :w], rt[:, 8:8 + c, :, :w], d).backward(torch.from_numpy(go).cuda().bfloat16())
    gl, gr = oracle.inner_product_volume_bwd(go, lf[:, 8:8 + c, :, :w], rf[:, 8:8 + c, :, :w])
    _check(lt.grad[:, 8:8 + c, :, :w], gl, "bf16", d, c, False)
    _check(rt.grad[:, 8:8 + c, :, :w], gr, "bf16", d, c, False)
    assert float(lt.grad[:, :8].abs().max()) == 0.0 and float(lt.grad[..., w:].abs().max()) == 0.0
    # D > 64, W % 8 != 0, C % 16 != 0: the SIMT kernels
    for shape in [(1, 32, 2, 200, 70), (1, 32, 2, 130, 24), (1, 24, 2, 136, 24)]:
        n, c, h, w, d = shape
        l, r, go = _case(shape, "bf16")
        gl, gr = oracle.inner_product_volume_bwd(go, l, r)
        lt = torch.from_numpy(l).cuda().bfloat16().requires_grad_(True)
        rt = torch.from_numpy(r).cuda().bfloat16().requires_grad_(True)
        rsm.inner_product_volume(lt, rt, d).backward(torch.from_numpy(go).cuda().bfloat16())
        _check(lt.grad, gl, "bf16", d, c, False)
        _check(rt.grad, gr, "bf16", d, c, False)
