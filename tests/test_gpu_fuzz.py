"""GPU suite: seeded random-shape sweep over every dispatch branch of the volume / regression kernels.

The kernels pick specialised paths from alignment and divisibility (16-byte rows, D % 8, C % 16 / 32, narrow
groups, TMA-addressable strides, ...).  Each case draws a shape that is deliberately on or just off those
boundaries, runs forward (all dtypes) and backward (fp32) through the public API and compares with the oracle."""
import numpy as np
import pytest
import torch

import oracle
from golden_io import round_to
from tolerances import GRAD_RTOL, RTOL_16, corr_atol_fp32, soft_argmax_atol

pytestmark = pytest.mark.gpu
DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()
    return m


def dev(a, dn="fp32", grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(DT[dn]).requires_grad_(grad)


def host(t):
    return t.detach().float().cpu().numpy()


def draw(seed):
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.choice([8, 16, 24, 64, 128, 136, 156, 240, 264]) + rng.choice([0, 0, 0, 1, 2, 4]))
    d = int(rng.choice([1, 8, 16, 24, 40, 48, 64, 70]) + rng.choice([0, 0, 1]))
    g = int(rng.choice([1, 2, 4, 8]))
    cpg = int(rng.choice([1, 2, 4, 8, 16, 32]))
    c = g * cpg
    if c > 64:
        c, g = 64, max(1, 64 // cpg)
    n, h = int(rng.integers(1, 3)), int(rng.integers(1, 4))
    dn = str(rng.choice(["fp32", "fp32", "bf16", "fp16"]))
    return rng, n, c, h, w, d, g, dn


@pytest.mark.parametrize("seed", range(48))
def test_random_shapes_volumes(rsm, seed):
    rng, n, c, h, w, d, g, dn = draw(seed)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    L, R = dev(l, dn), dev(r, dn)
    tag = f"n{n} c{c} h{h} w{w} d{d} g{g} {dn}"
    np.testing.assert_array_equal(host(rsm.concat_volume(L, R, d)), oracle.concat_volume(l, r, d), err_msg=tag)
    np.testing.assert_array_equal(host(rsm.interweave(L, R)), oracle.interweave(l, r), err_msg=tag)
    np.testing.assert_array_equal(host(rsm.shift_interweave_volume(L, R, d)), oracle.shift_interweave_volume(l, r, d), err_msg=tag)
    diff = host(rsm.difference_volume(L, R, d))
    np.testing.assert_allclose(diff, round_to(oracle.difference_volume(l, r, d), dn), rtol=0, atol=0, err_msg=tag)
    atol = corr_atol_fp32(c, np.abs(l).max(), np.abs(r).max())
    rt = 0.0 if dn == "fp32" else RTOL_16[dn]
    for mean in (False, True):
        ref = oracle.inner_product_volume(l, r, d, mean=mean, out_dtype=np.float32)
        np.testing.assert_allclose(host(rsm.inner_product_volume(L, R, d, mean=mean)), ref, rtol=rt, atol=atol * (4 if rt else 1), err_msg=tag)
        soft, amin, amax = rsm.inner_product_regress(L * 0.5, R * 0.5, d, mean=mean)
        vol = oracle.inner_product_volume(l * 0.5, r * 0.5, d, mean=mean, out_dtype=np.float32)
        np.testing.assert_allclose(host(soft), oracle.soft_argmax(vol), rtol=0, atol=soft_argmax_atol(d) * 4, err_msg=tag)
        assert (amin.cpu().numpy() != oracle.hard_argmin(vol)).mean() < 5e-3, tag
    gw = oracle.groupwise_volume(l, r, g, d, out_dtype=np.float32)
    np.testing.assert_allclose(host(rsm.groupwise_volume(L, R, g, d)), gw, rtol=rt, atol=atol * (4 if rt else 1), err_msg=tag)


@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_gradients(rsm, seed):
    rng, n, c, h, w, d, g, _ = draw(100 + seed)
    l = rng.standard_normal((n, c, h, w)).astype(np.float32)
    r = rng.standard_normal((n, c, h, w)).astype(np.float32)
    tag = f"n{n} c{c} h{h} w{w} d{d} g{g}"
    gs = GRAD_RTOL * np.sqrt(max(d, 1)) * 16
    cases = (
        (lambda a, b: rsm.concat_volume(a, b, d), lambda go: oracle.concat_volume_bwd(go), (n, 2 * c, h, w, d)),
        (lambda a, b: rsm.difference_volume(a, b, d), lambda go: oracle.difference_volume_bwd(go), (n, c, d, h, w)),
        (lambda a, b: rsm.inner_product_volume(a, b, d, mean=True), lambda go: oracle.inner_product_volume_bwd(go, l, r, mean=True), (n, d, h, w)),
        (lambda a, b: rsm.groupwise_volume(a, b, g, d), lambda go: oracle.groupwise_volume_bwd(go, l, r, g), (n, g, h, w, d)),
        (lambda a, b: rsm.interweave(a, b), lambda go: oracle.interweave_bwd(go), (n, 2 * c, h, w)),
        (lambda a, b: rsm.shift_interweave_volume(a, b, d), lambda go: oracle.shift_interweave_volume_bwd(go), (d, n, 2 * c, h, w)),
    )
    for fn, ofn, oshape in cases:
        gout = rng.standard_normal(oshape).astype(np.float32)
        L, R = dev(l, grad=True), dev(r, grad=True)
        fn(L, R).backward(dev(gout))
        gl, gr = ofn(gout)
        np.testing.assert_allclose(host(L.grad), gl, rtol=0, atol=gs, err_msg=tag)
        np.testing.assert_allclose(host(R.grad), gr, rtol=0, atol=gs, err_msg=tag)


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes_regression(rsm, seed):
    rng = np.random.default_rng(3000 + seed)
    n, d = int(rng.integers(1, 3)), int(rng.choice([1, 2, 7, 8, 9, 24, 48, 65, 192]))
    h, w = int(rng.integers(1, 9)), int(rng.choice([1, 3, 4, 8, 31, 32, 60]))
    dn = str(rng.choice(["fp32", "fp32", "bf16", "fp16"]))
    cost = round_to((rng.standard_normal((n, d, h, w)) * 4).astype(np.float32), dn)
    soft, amin, amax = rsm.regress(dev(cost, dn))
    np.testing.assert_array_equal(amin.cpu().numpy(), oracle.hard_argmin(cost))
    np.testing.assert_array_equal(amax.cpu().numpy(), oracle.hard_argmax(cost))
    atol = soft_argmax_atol(d) if dn == "fp32" else RTOL_16[dn] * d
    np.testing.assert_allclose(host(soft), oracle.soft_argmax(cost), rtol=0, atol=atol)
    if dn == "fp32":
        c = dev(cost, grad=True)
        gout = rng.standard_normal((n, h, w)).astype(np.float32)
        rsm.soft_argmax(c).backward(dev(gout))
        np.testing.assert_allclose(host(c.grad), oracle.soft_argmax_bwd(gout, cost), rtol=0, atol=GRAD_RTOL * d)


@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_prepost_and_loss(rsm, seed):
    """Pre / post steps and the loss on random frame sizes, alignments and resize ratios:
    aligned and unaligned rows, up- and down-sampling, crops, against the oracle."""
    rng = np.random.default_rng(5000 + seed)
    n, c = int(rng.integers(1, 3)), int(rng.choice([1, 3]))
    h, w = int(rng.integers(3, 40)), int(rng.integers(3, 70))
    align = int(rng.choice([1, 4, 8, 16, 64]))
    img = (rng.random((n, c, h, w)) * 255).astype(np.float32)
    tag = f"n{n} c{c} h{h} w{w} align{align}"
    x = dev(img, grad=True)
    out = rsm.prepare_input(x, align)
    ref = oracle.prepare_input(img, align, device_div=True)
    np.testing.assert_array_equal(host(out), ref, err_msg=tag)
    gout = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(dev(gout))
    np.testing.assert_array_equal(host(x.grad), oracle.prepare_input_bwd(gout, (h, w), device_div=True), err_msg=tag)

    hp, wp = ref.shape[2:]
    hs, ws = int(rng.integers(1, hp + 6)), int(rng.integers(1, wp + 6))
    disp = (rng.standard_normal((n, 1, hs, ws)) * 7).astype(np.float32)
    for mode in ("nearest", "bilinear"):
        t = dev(disp, grad=True)
        fin = rsm.finalize_disparity(t, (hp, wp), (h, w), mode=mode)
        want = oracle.finalize_disparity(disp, (hp, wp), (h, w), mode)
        if mode == "nearest":
            np.testing.assert_array_equal(host(fin), want, err_msg=tag)
        else:
            np.testing.assert_allclose(host(fin), want, atol=1e-5, rtol=1e-5, err_msg=tag)
        go = rng.standard_normal((n, 1, h, w)).astype(np.float32)
        fin.backward(dev(go))
        np.testing.assert_allclose(host(t.grad), oracle.finalize_disparity_bwd(go, disp.shape, (hp, wp), mode),
                                   atol=1e-4, rtol=3e-4, err_msg=f"{tag} {mode} {hs}x{ws}")

    gt = (rng.standard_normal((n, 1, h, w)) * 30).astype(np.float32)
    valid = (rng.random((n, h, w)) > 0.25).astype(np.float32)
    sizes = [(int(rng.integers(1, h + 3)), int(rng.integers(1, w + 3))), (h, w)]
    preds = [(rng.standard_normal((n, 1) + s) * 10).astype(np.float32) for s in sizes]
    tp = [dev(p, grad=True) for p in preds]
    loss = rsm.SequenceLoss(0.85, 40.0)(tp, dev(gt), dev(valid))
    np.testing.assert_allclose(host(loss), oracle.sequence_loss(preds, gt, valid, 0.85, 40.0), rtol=3e-6, err_msg=tag)
    loss.backward()
    for p, gp in zip(tp, oracle.sequence_loss_bwd(preds, gt, valid, 0.85, 40.0)):
        np.testing.assert_allclose(host(p.grad), gp, atol=1e-8, rtol=1e-5, err_msg=tag)
    got = rsm.get_flow_map_metrics(dev(gt), tp[-1].detach(), dev(valid))
    want = oracle.flow_map_metrics(gt, preds[-1], valid)
    for key in want:
        np.testing.assert_allclose(got[key], want[key], atol=1e-7, rtol=3e-6, err_msg=f"{tag} {key}")
