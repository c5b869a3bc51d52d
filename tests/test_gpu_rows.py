"""GPU parity of the row-streaming fused inner-product -> regression kernel (csrc/rsm_corr_rows.cu, 16-bit features):
soft-argmax within 1e-4*D px of the oracle on the same rounded inputs, argmin / argmax bit-exact on dyadic inputs
(exact sums; ties take the first index) and within 0.2 % flips on random ones, mean (exact division for non-power-of-two
C) and sum, one and two accumulator groups, windows that wrap the atom ring, rows narrower than a tile, D = 1 and the
largest D, NaNs inside the band and inside the x < d fill region, and the fall-backs for shapes it does not cover.
Reference: make_correlation_volume model/mobile_disp_net_c.py:188-205, TorchInnerProductCost
cost_volume/inner_product.py:11-42, followed by the soft-argmax of model/mobile_stereo_net.py:144-147 / torch.argmin."""
import numpy as np
import pytest
import torch

import oracle
from golden_io import round_to
from tolerances import soft_argmax_atol

pytestmark = pytest.mark.gpu

DT = {"fp16": torch.float16, "bf16": torch.bfloat16}
SHAPES = [(1, 16, 2, 128, 16), (1, 16, 3, 240, 48), (2, 64, 5, 240, 48), (1, 32, 4, 312, 48), (1, 64, 2, 480, 128),
          (1, 16, 2, 72, 19), (1, 128, 2, 480, 192), (1, 32, 3, 200, 130), (2, 16, 2, 304, 260), (1, 48, 2, 136, 1),
          (1, 16, 2, 8, 24), (1, 96, 3, 264, 65), (1, 16, 1, 520, 384), (3, 32, 7, 96, 64), (1, 112, 2, 1000, 200)]


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()
    return m


def _pair(shape, dn, seed=1, scale=0.5):
    rng = np.random.default_rng(seed)
    l = round_to(rng.standard_normal(shape).astype(np.float32) * scale, dn)
    r = round_to(rng.standard_normal(shape).astype(np.float32) * scale, dn)
    return l, r, torch.from_numpy(l).cuda().to(DT[dn]), torch.from_numpy(r).cuda().to(DT[dn])


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dn", ["bf16", "fp16"])
@pytest.mark.parametrize("mean", [False, True])
def test_rows_random(rsm, shape, dn, mean):
    n, c, h, w, d = shape
    l, r, lt, rt = _pair((n, c, h, w), dn)
    vol = oracle.inner_product_volume(l, r, d, mean=mean)
    want = oracle.soft_argmax(vol)
    soft, amin, amax = rsm.inner_product_regress(lt, rt, d, mean=mean)
    soft_only, _, _ = rsm.inner_product_regress(lt, rt, d, mean=mean, argmin=False, argmax=False)
    assert np.abs(soft.cpu().numpy() - want).max() <= soft_argmax_atol(d)
    assert np.abs(soft_only.cpu().numpy() - want).max() <= soft_argmax_atol(d)
    # the accumulation order differs from the oracle's: near-ties may flip, nothing else
    assert (amin.cpu().numpy() != oracle.hard_argmin(vol)).mean() < 2e-3
    assert (amax.cpu().numpy() != oracle.hard_argmax(vol)).mean() < 2e-3


@pytest.mark.parametrize("shape,mean", [((2, 32, 6, 160, 24), False), ((1, 64, 3, 240, 48), True), ((1, 128, 2, 480, 192), False),
                                        ((1, 16, 2, 320, 100), True), ((1, 48, 3, 200, 70), True), ((2, 16, 2, 520, 384), False)])
def test_rows_dyadic_bit_exact(rsm, shape, mean):
    n, c, h, w, d = shape
    rng = np.random.default_rng(7)
    l = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    r = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    vol = oracle.inner_product_volume(l, r, d, mean=mean)
    soft, amin, amax = rsm.inner_product_regress(torch.from_numpy(l).cuda().bfloat16(), torch.from_numpy(r).cuda().bfloat16(), d, mean=mean)
    np.testing.assert_array_equal(amin.cpu().numpy(), oracle.hard_argmin(vol))
    np.testing.assert_array_equal(amax.cpu().numpy(), oracle.hard_argmax(vol))
    assert np.abs(soft.cpu().numpy() - oracle.soft_argmax(vol)).max() <= soft_argmax_atol(d)


def test_rows_nan(rsm):
    """A NaN left feature at x = 5 poisons disparities 0..5 of that pixel only (the reference never computes x < d); a NaN
    right feature at x' = 100 poisons (x, d) with x - d = 100."""
    n, c, h, w, d = 1, 16, 2, 200, 48
    l, r, _, _ = _pair((n, c, h, w), "bf16", seed=3, scale=1.0)
    l[0, 3, 0, 5] = np.nan
    r[0, 2, 1, 100] = np.nan
    vol = oracle.inner_product_volume(l, r, d)
    soft, amin, amax = rsm.inner_product_regress(torch.from_numpy(l).cuda().bfloat16(), torch.from_numpy(r).cuda().bfloat16(), d)
    np.testing.assert_array_equal(amin.cpu().numpy(), oracle.hard_argmin(vol))
    np.testing.assert_array_equal(amax.cpu().numpy(), oracle.hard_argmax(vol))
    np.testing.assert_array_equal(np.isnan(soft.cpu().numpy()), np.isnan(oracle.soft_argmax(vol)))


def test_rows_strided_views_and_fallbacks(rsm):
    """Width-cropped / channel-sliced views go through the tensor maps as they are; shapes the kernel does not cover
    (W % 8 != 0: TMA strides, C > 128, C % 16 != 0, D > 384) fall back to the other fused kernels with the same results."""
    n, c, h, w, d = 2, 32, 5, 248, 48
    l, r, lt, rt = _pair((n, c + 16, h, w + 8), "bf16")
    lv, rv = lt[:, 8:8 + c, :, :w], rt[:, 8:8 + c, :, :w]
    vol = oracle.inner_product_volume(l[:, 8:8 + c, :, :w], r[:, 8:8 + c, :, :w], d)
    soft, amin, _ = rsm.inner_product_regress(lv, rv, d)
    assert np.abs(soft.cpu().numpy() - oracle.soft_argmax(vol)).max() <= soft_argmax_atol(d)
    assert (amin.cpu().numpy() != oracle.hard_argmin(vol)).mean() < 2e-3
    for shape in [(1, 32, 3, 250, 48), (1, 144, 2, 136, 24), (1, 24, 2, 136, 24), (1, 16, 1, 520, 400)]:
        n, c, h, w, d = shape
        l, r, lt, rt = _pair((n, c, h, w), "bf16")
        vol = oracle.inner_product_volume(l, r, d)
        soft, amin, _ = rsm.inner_product_regress(lt, rt, d)
        assert np.abs(soft.cpu().numpy() - oracle.soft_argmax(vol)).max() <= soft_argmax_atol(d)
        assert (amin.cpu().numpy() != oracle.hard_argmin(vol)).mean() < 2e-3


def test_rows_full_size_matches_volume_path(rsm):
    """cfg2 slice at full width/height (N = 4): the fused result against the materialised tcgen05 volume + regression."""
    n, c, h, w, d = 4, 64, 144, 240, 48
    lt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    rt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    vol = rsm.make_correlation_volume(lt, rt, d).float()
    want = (torch.softmax(vol, 1) * torch.arange(d, device="cuda", dtype=torch.float32).view(1, d, 1, 1)).sum(1)
    soft, amin, amax = rsm.inner_product_regress(lt, rt, d, mean=True)
    # the volume path rounds the volume to bf16 before the softmax; the fused path keeps fp32 accumulators
    assert (soft - want).abs().max().item() < 0.25
    assert (soft - want).abs().mean().item() < 5e-3
