"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI of librsm_b200.so, against
(1) the committed golden vectors produced by executing the reference, (2) the numpy oracle on
seeded inputs, (3) size-independent properties at BASELINE.json's full sizes.

Bars (tests/tolerances.py): bit-exact for concatenate / interweave / difference, for hard
argmin/argmax on a given volume and for every reduction on dyadic inputs; stated fp32 / 16-bit
tolerances elsewhere.
"""
import numpy as np
import pytest
import torch

import oracle
from golden_io import load, names, round_to
from tolerances import (GRAD_RTOL, RTOL_16, V4_TAIL_ATOL, corr_atol_fp32, soft_argmax_atol)

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()  # fail loudly if the extension is missing
    return m


def dev(a, dname="fp32", grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(DT[dname])
    return t.requires_grad_(grad)


def host(t):
    return t.detach().float().cpu().numpy()


def close(a, b, atol, rtol=0.0):
    np.testing.assert_allclose(host(a) if torch.is_tensor(a) else a, b, atol=atol, rtol=rtol)


def equal(a, b):
    if torch.is_tensor(a):
        a = a.cpu().numpy() if a.dtype == torch.int64 else host(a)
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------ golden vectors: volumes
@pytest.mark.parametrize("name", names("vol_"))
def test_volume_goldens(rsm, name):
    g, m = load(name)
    dn, d, ng, c = m["dtype"], m["D"], m["G"], m["C"]
    exact = m["kind"] == "dyadic"
    lmax, rmax = np.abs(g["left"]).max(), np.abs(g["right"]).max()

    def fresh():
        return dev(g["left"], dn, True), dev(g["right"], dn, True)

    # --- pure data movement: bit exact, forward and backward
    for op, fn in (("concat", lambda a, b: rsm.TorchConcatenateCost(d)(a, b)),
                   ("interweave", lambda a, b: rsm.TorchInterweaveCost()(a, b)),
                   ("interweave_v4", rsm.interweave_tensors),
                   ("difference", lambda a, b: rsm.make_cost_volume(a, b, d))):
        l, r = fresh()
        out = fn(l, r)
        assert out.dtype == DT[dn] and out.is_contiguous()
        equal(out, g[f"{op}.out"])
        out.backward(dev(g[f"{op}.gout"], dn))
        if dn == "fp32":
            tol = 0.0 if op.startswith("interweave") else GRAD_RTOL * np.sqrt(d)
            close(l.grad, g[f"{op}.gleft"], tol)
            close(r.grad, g[f"{op}.gright"], tol)
    # --- reductions
    if dn == "fp32":
        atol, rtol = (0.0 if exact else corr_atol_fp32(c, lmax, rmax)), 0.0
    else:
        atol, rtol = RTOL_16[dn] * np.sqrt(c) * lmax * rmax, RTOL_16[dn] * 2
    gs = GRAD_RTOL * np.sqrt(d) * max(lmax, rmax) * 4
    for op, fn in (("inner", lambda a, b: rsm.TorchInnerProductCost(d)(a, b)),
                   ("corr_mean", lambda a, b: rsm.make_correlation_volume(a, b, d)),
                   ("groupwise", lambda a, b: rsm.TorchGroupwiseCost(ng, d, out_dtype=torch.float32)(a, b))):
        l, r = fresh()
        out = fn(l, r)
        close(out, g[f"{op}.out"], atol, rtol)
        if dn == "fp32":
            out.backward(dev(g[f"{op}.gout"]))
            close(l.grad, g[f"{op}.gleft"], gs)
            close(r.grad, g[f"{op}.gright"], gs)


def test_noncontiguous_slices(rsm):
    """v4 passes width-cropped non-contiguous views (mobile_stereo_net_v4.py:446)."""
    g, m = load("noncontig_interweave")
    i = m["i"]
    fl, fr = dev(g["featL"]), dev(g["featR"])
    a, b = fl[:, :, :, i:], fr[:, :, :, :-i]
    assert not a.is_contiguous()
    equal(rsm.interweave_tensors(a, b), g["out"])
    equal(rsm.TorchConcatenateCost(5)(a, b), g["concat"])
    close(rsm.TorchInnerProductCost(5)(a, b), g["inner"], 1e-5)


# --------------------------------------------------------------- golden vectors: regression
@pytest.mark.parametrize("name", names("regress_"))
def test_regression_goldens(rsm, name):
    g, m = load(name)
    dn = m.get("dtype", "fp32")
    cost = dev(g["cost"], dn, "e" in g and dn == "fp32")
    soft, amin, amax = rsm.regress(cost)
    equal(amin, g["argmin"])
    equal(amax, g["argmax"])
    equal(rsm.hard_argmin(cost), g["argmin"])
    equal(rsm.hard_argmax(cost), g["argmax"])
    if "e" not in g:
        return
    atol = soft_argmax_atol(m["D"]) if dn == "fp32" else RTOL_16[dn] * m["D"]
    close(soft, g["e"], atol)
    close(rsm.disparity_regression_dispnetc(cost, m["D"]), g["e_keepdim"], atol)
    close(rsm.disparity_regression_v4(torch.softmax(cost.detach().float(), 1).to(DT[dn]), m["D"]), g["e"],
          atol * (4 if dn != "fp32" else 1))
    if dn == "fp32":
        rsm.disparity_regression_dispnetc(cost, m["D"]).backward(dev(g["gout"]))
        close(cost.grad, g["gcost"], GRAD_RTOL * m["D"])


@pytest.mark.parametrize("name", names("tail_"))
def test_v4_tail_goldens(rsm, name):
    g, m = load(name)
    cost = dev(g["cost"], "fp32", True)
    pred, amin, amax = rsm.upsample_regress(cost, m["D"], m["H"], m["W"], argmin=True, argmax=True)
    close(pred, g["pred"], V4_TAIL_ATOL)
    pred.backward(dev(g["gout"]))
    close(cost.grad, g["gcost"], GRAD_RTOL * m["D"])
    # hard extrema over the upsampled volume: identical wherever the top-2 gap is resolvable
    fine = g["fine"]
    srt = np.sort(fine, axis=1)
    ok_max = (srt[:, -1] - srt[:, -2]) > 1e-4
    ok_min = (srt[:, 1] - srt[:, 0]) > 1e-4
    assert np.array_equal(amax.cpu().numpy()[ok_max], g["argmax"][ok_max])
    assert np.array_equal(amin.cpu().numpy()[ok_min], g["argmin"][ok_min])


# ------------------------------------------------------------- golden vectors: call sites
def test_callsite_v1(rsm):
    g, m = load("callsite_v1")
    equal(rsm.make_cost_volume(dev(g["lf"]), dev(g["rf"]), m["max_disp"]), g["volume"])
    close(rsm.softmax_regression(dev(g["filtered"])), g["regressed"], soft_argmax_atol(m["max_disp"]))


def test_callsite_dispnetc(rsm):
    g, m = load("callsite_dispnetc")
    c = g["lf"].shape[1]
    atol = corr_atol_fp32(c, np.abs(g["lf"]).max(), np.abs(g["rf"]).max())
    close(rsm.make_correlation_volume(dev(g["lf"]), dev(g["rf"]), m["max_disp"]), g["volume"], atol)


def test_callsite_v4(rsm):
    g, m = load("callsite_v4")
    for k in range(m["n_iw"]):
        equal(rsm.interweave_tensors(dev(g[f"iw{k}.a"]), dev(g[f"iw{k}.b"])), g[f"iw{k}.out"])
    pred = rsm.v4_head(dev(g["cost3"]), m["maxdisp"], m["H"], m["W"])
    close(-pred.unsqueeze(1), g["final"], V4_TAIL_ATOL)


# ------------------------------------------------------------------ oracle on seeded inputs
SHAPES = [  # N, C, H, W, D, G
    (2, 32, 12, 156, 24, 8),     # cfg1-like feature shape
    (1, 64, 9, 240, 48, 16),     # cfg2-like row
    (1, 32, 6, 312, 48, 8),      # cfg3-like row
    (1, 12, 5, 67, 19, 3),       # nothing divisible by 4
    (1, 20, 3, 130, 70, 5),      # D > 64: two disparity chunks
    (3, 5, 4, 9, 13, 5),         # D > W
    (1, 40, 3, 132, 70, 5),      # 16-byte aligned rows, two disparity chunks, ragged channel block (8x8 / 8xDT tiles)
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dn", ["fp32", "bf16", "fp16"])
def test_volumes_vs_oracle(rsm, shape, dn):
    n, c, h, w, d, ng = shape
    rng = np.random.default_rng(1234)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    L, R = dev(l, dn), dev(r, dn)
    equal(rsm.concat_volume(L, R, d), oracle.concat_volume(l, r, d))
    equal(rsm.interweave(L, R), oracle.interweave(l, r))
    equal(rsm.difference_volume(L, R, d), round_to(oracle.difference_volume(l, r, d), dn))
    lmax, rmax = np.abs(l).max(), np.abs(r).max()
    if dn == "fp32":
        atol, rtol = corr_atol_fp32(c, lmax, rmax), 0.0
    else:   # fp32 oracle on the same rounded inputs; only the output cast differs
        atol, rtol = 1e-5 * np.sqrt(c) * lmax * rmax, RTOL_16[dn]
    close(rsm.inner_product_volume(L, R, d), oracle.inner_product_volume(l, r, d, out_dtype=np.float32), atol, rtol)
    close(rsm.inner_product_volume(L, R, d, mean=True),
          oracle.inner_product_volume(l, r, d, mean=True, out_dtype=np.float32), atol, rtol)
    close(rsm.groupwise_volume(L, R, ng, d, out_dtype=torch.float32), oracle.groupwise_volume(l, r, ng, d),
          corr_atol_fp32(c // ng, lmax, rmax))


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_dyadic_bit_exact(rsm, shape):
    """k/8 inputs: fp32 sums are exact, so every reduction must match the oracle bit for bit,
    and so must the fused build + argmin."""
    n, c, h, w, d, ng = shape
    rng = np.random.default_rng(7)
    l = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    r = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    L, R = dev(l), dev(r)
    vol = oracle.inner_product_volume(l, r, d)
    equal(rsm.inner_product_volume(L, R, d), vol)
    equal(rsm.groupwise_volume(L, R, ng, d), oracle.groupwise_volume(l, r, ng, d))
    soft, amin, amax = rsm.inner_product_regress(L, R, d)
    equal(amin, oracle.hard_argmin(vol))
    equal(amax, oracle.hard_argmax(vol))
    close(soft, oracle.soft_argmax(vol), soft_argmax_atol(d))


@pytest.mark.parametrize("shape", SHAPES)
def test_volume_grads_vs_oracle(rsm, shape):
    n, c, h, w, d, ng = shape
    rng = np.random.default_rng(99)
    l = rng.standard_normal((n, c, h, w)).astype(np.float32)
    r = rng.standard_normal((n, c, h, w)).astype(np.float32)
    gs = GRAD_RTOL * np.sqrt(d) * 16
    for fn, ofn, oshape in (
        (lambda a, b: rsm.concat_volume(a, b, d), lambda g_: oracle.concat_volume_bwd(g_), (n, 2 * c, h, w, d)),
        (lambda a, b: rsm.difference_volume(a, b, d), lambda g_: oracle.difference_volume_bwd(g_), (n, c, d, h, w)),
        (lambda a, b: rsm.inner_product_volume(a, b, d), lambda g_: oracle.inner_product_volume_bwd(g_, l, r), (n, d, h, w)),
        (lambda a, b: rsm.inner_product_volume(a, b, d, mean=True),
         lambda g_: oracle.inner_product_volume_bwd(g_, l, r, mean=True), (n, d, h, w)),
        (lambda a, b: rsm.groupwise_volume(a, b, ng, d), lambda g_: oracle.groupwise_volume_bwd(g_, l, r, ng),
         (n, ng, h, w, d)),
    ):
        gout = rng.standard_normal(oshape).astype(np.float32)
        L, R = dev(l, grad=True), dev(r, grad=True)
        fn(L, R).backward(dev(gout))
        gl, gr = ofn(gout)
        close(L.grad, gl, gs)
        close(R.grad, gr, gs)


@pytest.mark.parametrize("shape", [(2, 24, 12, 40), (1, 192, 8, 60), (1, 48, 5, 33), (1, 7, 3, 5), (2, 1, 2, 2)])
@pytest.mark.parametrize("dn", ["fp32", "bf16", "fp16"])
def test_regress_vs_oracle(rsm, shape, dn):
    n, d, h, w = shape
    rng = np.random.default_rng(5)
    cost = round_to((rng.standard_normal(shape) * 4).astype(np.float32), dn)
    soft, amin, amax = rsm.regress(dev(cost, dn))
    equal(amin, oracle.hard_argmin(cost))     # bit exact on the given volume
    equal(amax, oracle.hard_argmax(cost))
    atol = soft_argmax_atol(d) if dn == "fp32" else RTOL_16[dn] * d
    close(soft, oracle.soft_argmax(cost), atol)
    if dn == "fp32":
        c = dev(cost, grad=True)
        gout = rng.standard_normal((n, h, w)).astype(np.float32)
        rsm.soft_argmax(c).backward(dev(gout))
        close(c.grad, oracle.soft_argmax_bwd(gout, cost), GRAD_RTOL * d)


def test_regress_ties_everywhere(rsm):
    """integer-valued costs: massive ties, first occurrence must win at every pixel."""
    rng = np.random.default_rng(3)
    cost = rng.integers(-2, 3, (2, 37, 9, 20)).astype(np.float32)
    _, amin, amax = rsm.regress(dev(cost))
    equal(amin, oracle.hard_argmin(cost))
    equal(amax, oracle.hard_argmax(cost))


@pytest.mark.parametrize("geom", [(1, 48, 6, 10, 192, 24, 40), (2, 12, 5, 7, 48, 20, 28), (1, 5, 3, 4, 13, 7, 10),
                                  (1, 9, 8, 8, 9, 4, 4), (2, 11, 23, 37, 44, 92, 148), (1, 8, 4, 10, 32, 16, 40)])
def test_v4_tail_vs_oracle(rsm, geom):
    b, dc, hc, wc, d, h, w = geom
    rng = np.random.default_rng(11)
    cost = (rng.standard_normal((b, dc, hc, wc)) * 3).astype(np.float32)
    c = dev(cost, grad=True)
    pred = rsm.upsample_regress(c, d, h, w)
    close(pred, oracle.v4_tail(cost, d, h, w), V4_TAIL_ATOL)
    gout = rng.standard_normal((b, h, w)).astype(np.float32)
    pred.backward(dev(gout))
    close(c.grad, oracle.v4_tail_bwd(gout, cost, d, h, w), GRAD_RTOL * d)


@pytest.mark.parametrize("shape", [(1, 16, 4, 100, 48), (2, 64, 3, 70, 192), (1, 8, 2, 33, 24)])
@pytest.mark.parametrize("mean", [False, True])
def test_fused_inner_regress_vs_oracle(rsm, shape, mean):
    n, c, h, w, d = shape
    rng = np.random.default_rng(21)
    l = rng.standard_normal((n, c, h, w)).astype(np.float32)
    r = rng.standard_normal((n, c, h, w)).astype(np.float32)
    soft, amin, amax = rsm.inner_product_regress(dev(l), dev(r), d, mean=mean)
    vol = oracle.inner_product_volume(l, r, d, mean=mean)
    close(soft, oracle.soft_argmax(vol), soft_argmax_atol(d) * 4)
    # accumulation order differs from the oracle, so near-ties may flip: report, bound at 0.1 %
    mism = (amax.cpu().numpy() != oracle.hard_argmax(vol)).mean()
    assert mism < 1e-3, f"fused argmax mismatch rate {mism}"


# -------------------------------------------------- full-size properties (BASELINE configs)
def test_fullsize_concat_groupwise_properties(rsm):
    """cfg3 shapes per GPU (N=2 here to bound memory): every disparity slice of the volume equals
    the shifted inputs; checked with torch ops on the device."""
    n, c, h, w, d, ng = 2, 32, 96, 312, 48, 8
    gen = torch.Generator(device="cuda").manual_seed(1234)
    L = torch.randn((n, c, h, w), device="cuda", generator=gen)
    R = torch.randn((n, c, h, w), device="cuda", generator=gen)
    vol = rsm.concat_volume(L, R, d)
    gw = rsm.groupwise_volume(L, R, ng, d)
    for dd in (0, 1, 17, 47):
        assert torch.equal(vol[:, :c, :, dd:, dd], L[:, :, :, dd:])
        assert torch.equal(vol[:, c:, :, dd:, dd], R[:, :, :, : w - dd])
        assert vol[:, :, :, :dd, dd].abs().sum().item() == 0.0
        ref = (L[:, :, :, dd:] * R[:, :, :, : w - dd]).view(n, ng, c // ng, h, w - dd).mean(2)
        torch.testing.assert_close(gw[:, :, :, dd:, dd], ref, atol=2e-5, rtol=0)
        assert gw[:, :, :, :dd, dd].abs().sum().item() == 0.0
    # checksum of checksums: sum over d of the left half = L * min(x+1, D)
    cnt = torch.clamp(torch.arange(w, device="cuda") + 1, max=d).float()
    torch.testing.assert_close(vol[:, :c].sum(-1), L * cnt, atol=1e-4, rtol=1e-5)


def test_fullsize_correlation_cfg2(rsm):
    """cfg2: (N,64,144,240), D=48 mean correlation (N=4 here), against torch ops on the device."""
    n, c, h, w, d = 4, 64, 144, 240, 48
    gen = torch.Generator(device="cuda").manual_seed(1234)
    L = torch.randn((n, c, h, w), device="cuda", generator=gen)
    R = torch.randn((n, c, h, w), device="cuda", generator=gen)
    vol = rsm.make_correlation_volume(L, R, d)
    for dd in (0, 5, 47):
        ref = (L[:, :, :, dd:] * R[:, :, :, : w - dd]).mean(1)
        torch.testing.assert_close(vol[:, dd, :, dd:], ref, atol=2e-5, rtol=0)
        assert vol[:, dd, :, :dd].abs().sum().item() == 0.0


def test_fullsize_regression_cfg5(rsm):
    """cfg5: one (192,1080,1920) fp32 volume: argmin/argmax bit-exact vs torch on the device,
    soft-argmax within tolerance, invariance under a constant shift of the cost."""
    gen = torch.Generator(device="cuda").manual_seed(1234)
    cost = torch.randn((1, 192, 1080, 1920), device="cuda", generator=gen) * 4
    soft, amin, amax = rsm.regress(cost)
    assert torch.equal(amin, torch.argmin(cost, 1))
    assert torch.equal(amax, torch.argmax(cost, 1))
    ref = (torch.softmax(cost, 1) * torch.arange(192, device="cuda").view(1, -1, 1, 1)).sum(1)
    torch.testing.assert_close(soft, ref, atol=soft_argmax_atol(192), rtol=0)
    soft2 = rsm.soft_argmax(cost + 3.0)
    torch.testing.assert_close(soft2, soft, atol=soft_argmax_atol(192), rtol=0)


def test_fullsize_v4_tail_cfg3(rsm):
    """cfg3 tail: (2,48,96,312) -> (2,384,1248) against interpolate+softmax on the device."""
    gen = torch.Generator(device="cuda").manual_seed(1234)
    cost = torch.randn((2, 48, 96, 312), device="cuda", generator=gen) * 3
    pred = rsm.v4_head(cost, 192, 384, 1248)
    fine = torch.nn.functional.interpolate(cost.unsqueeze(1), [192, 384, 1248], mode="trilinear").squeeze(1)
    ref = (torch.softmax(fine, 1) * torch.arange(192, device="cuda").view(1, -1, 1, 1)).sum(1)
    torch.testing.assert_close(pred, ref, atol=V4_TAIL_ATOL, rtol=0)


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


def test_fullsize_adjoint_identities_cfg3(rsm):
    """cfg3 shapes, the whole per-GPU batch: every volume is (bi)linear in the features, so for any upstream gradient
    <gV, V(L, R)> = <gL, L> = <gR, R> (group-wise) and <gV, V> = <gL, L> + <gR, R> (concatenation) -- a
    size-independent check of the adjoint kernels against the forward kernels (which the tests above pin to torch)."""
    n, c, h, w, d, ng = 8, 32, 96, 312, 48, 8
    gen = torch.Generator(device="cuda").manual_seed(4321)
    L = torch.randn((n, c, h, w), device="cuda", generator=gen).requires_grad_(True)
    R = torch.randn((n, c, h, w), device="cuda", generator=gen).requires_grad_(True)
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1e-2)):
        l, r = L.detach().to(dt).requires_grad_(True), R.detach().to(dt).requires_grad_(True)
        vol = rsm.groupwise_volume(l, r, ng, d)
        gv = torch.randn(vol.shape, device="cuda", generator=gen).to(dt)
        gl, gr = torch.autograd.grad(vol, (l, r), gv)
        ref = _dot(gv, vol)
        scale = (gv.double().abs() * vol.double().abs()).sum().item()
        assert abs(_dot(gl, l) - ref) <= tol * scale, (dt, _dot(gl, l), ref)
        assert abs(_dot(gr, r) - ref) <= tol * scale, (dt, _dot(gr, r), ref)
        del vol, gv
        vol = rsm.concat_volume(l, r, d)
        gv = torch.randn(vol.shape, device="cuda", generator=gen).to(dt)
        gl, gr = torch.autograd.grad(vol, (l, r), gv)
        ref = _dot(gv, vol)
        scale = (gv.double().abs() * vol.double().abs()).sum().item()
        assert abs(_dot(gl, l) + _dot(gr, r) - ref) <= tol * scale, (dt, ref)
        del vol, gv


@pytest.mark.parametrize("shape", [(8, 64, 144, 240, 48), (1, 128, 270, 480, 192), (1, 64, 270, 480, 96)])
def test_fullsize_inner_adjoint_identity(rsm, shape):
    """cfg2 (DispNetC's correlation, 8 of the 32 pairs) and cfg4's large-D points: <gV, V> = <gL, L> = <gR, R> for the
    inner-product volume, fp32 (SIMT adjoint) and bf16 (tcgen05 adjoint; D = 96 / 192 as chunks of 64 disparities)."""
    n, c, h, w, d = shape
    gen = torch.Generator(device="cuda").manual_seed(99)
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1e-2)):
        l = torch.randn((n, c, h, w), device="cuda", generator=gen).to(dt).requires_grad_(True)
        r = torch.randn((n, c, h, w), device="cuda", generator=gen).to(dt).requires_grad_(True)
        vol = rsm.inner_product_volume(l, r, d, mean=True)
        gv = torch.randn(vol.shape, device="cuda", generator=gen).to(dt)
        gl, gr = torch.autograd.grad(vol, (l, r), gv)
        ref = _dot(gv, vol)
        scale = (gv.double().abs() * vol.double().abs()).sum().item()
        assert abs(_dot(gl, l) - ref) <= tol * scale, (dt, _dot(gl, l), ref)
        assert abs(_dot(gr, r) - ref) <= tol * scale, (dt, _dot(gr, r), ref)


def test_fullsize_v4_head_adjoint_cfg3(rsm):
    """cfg3 head, one pair at full size (48 x 96 x 312 -> 192 x 384 x 1248): the tile + combine adjoint against autograd
    through interpolate -> softmax -> expectation on the device."""
    gen = torch.Generator(device="cuda").manual_seed(77)
    cost = (torch.randn((1, 48, 96, 312), device="cuda", generator=gen) * 3)
    gout = torch.randn((1, 384, 1248), device="cuda", generator=gen)
    c = cost.clone().requires_grad_(True)
    rsm.v4_head(c, 192, 384, 1248).backward(gout)
    ref = cost.clone().requires_grad_(True)
    fine = torch.nn.functional.interpolate(ref.unsqueeze(1), [192, 384, 1248], mode="trilinear").squeeze(1)
    (torch.softmax(fine, 1) * torch.arange(192, device="cuda", dtype=torch.float32).view(1, -1, 1, 1)).sum(1).backward(gout)
    err = (c.grad - ref.grad).abs().max().item()
    assert err <= GRAD_RTOL * 192 * max(1.0, ref.grad.abs().max().item()), err


# ------------------------------------------------------------------------ error behaviour
def test_errors(rsm):
    x = torch.zeros((1, 6, 2, 8), device="cuda")
    with pytest.raises(AssertionError, match="groupwise cost channel"):   # groupwise.py:15-17
        rsm.TorchGroupwiseCost(4, 3)(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rsm.concat_volume(x.cpu(), x.cpu(), 2)
    with pytest.raises(TypeError):
        rsm.concat_volume(x.double(), x.double(), 2)
    with pytest.raises(RuntimeError):
        rsm.concat_volume(x, x[:, :3], 2)
    assert str(rsm.TorchInnerProductCost(4)) == "TorchInnerProductCost | aijk,aijh->ajkh"
    assert str(rsm.TorchInterweaveCost()) == "TorchInterweaveCost"
    assert rsm.concat_volume(x[:0], x[:0], 2).shape == (0, 12, 2, 8, 2)   # empty batch


def test_streams_and_autocast(rsm):
    """ops enqueue on the caller's current stream and accept autocast's fp16 features (F11)."""
    s = torch.cuda.Stream()
    L = torch.randn((1, 8, 4, 32), device="cuda")
    with torch.cuda.stream(s):
        with torch.autocast("cuda", dtype=torch.float16):
            v = rsm.inner_product_volume(L.half(), L.half(), 4)
    s.synchronize()
    assert v.dtype == torch.float16
    ref = oracle.inner_product_volume(host(L.half()), host(L.half()), 4, out_dtype=np.float32)
    close(v, ref, 1e-2, RTOL_16["fp16"])


def test_epe_delta_synthetic(rsm):
    """EPE of the soft-argmax disparity on a synthetic pair (right = left shifted by a known
    disparity) for the CUDA path vs the oracle path: the delta must be negligible."""
    rng = np.random.default_rng(0)
    n, c, h, w, d, true_d = 1, 16, 8, 96, 24, 7
    l = rng.standard_normal((n, c, h, w)).astype(np.float32)
    r = np.zeros_like(l)
    r[..., : w - true_d] = l[..., true_d:]
    vol = rsm.inner_product_volume(dev(l), dev(r), d)
    ours = host(rsm.soft_argmax(vol * 4.0))
    ref = oracle.soft_argmax(oracle.inner_product_volume(l, r, d) * 4.0)
    valid = np.zeros((n, h, w), bool)
    valid[..., d:] = True
    epe_ours = np.abs(ours - true_d)[valid].mean()
    epe_ref = np.abs(ref - true_d)[valid].mean()
    assert abs(epe_ours - epe_ref) < 1e-4, (epe_ours, epe_ref)
    assert epe_ours < 0.5


# ----------------------------------------------------------- tcgen05 (tensor-core) paths
@pytest.mark.parametrize("shape", [(1, 16, 3, 240, 48), (2, 64, 4, 240, 48), (1, 32, 2, 312, 48),
                                   (1, 128, 2, 480, 192), (1, 16, 2, 67, 19), (1, 48, 2, 130, 130)])
@pytest.mark.parametrize("dn", ["bf16", "fp16"])
def test_tcgen05_inner_vs_oracle_and_simt(rsm, shape, dn):
    """16-bit features with C % 16 == 0 take the tcgen05 banded-GEMM kernel: check it against the fp32
    oracle on the same rounded inputs AND against the SIMT fp32 kernel fed the same (rounded) values."""
    n, c, h, w, d = shape
    rng = np.random.default_rng(31)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    L, R = dev(l, dn), dev(r, dn)
    ref = oracle.inner_product_volume(l, r, d, mean=True, out_dtype=np.float32)
    tc32 = rsm.inner_product_volume(L, R, d, mean=True, out_dtype=torch.float32)
    tc16 = rsm.inner_product_volume(L, R, d, mean=True)
    simt32 = rsm.inner_product_volume(dev(l), dev(r), d, mean=True)
    atol = 1e-5 * np.sqrt(c) * np.abs(l).max() * np.abs(r).max()
    close(tc32, ref, atol)
    close(simt32, ref, atol)
    close(tc16, ref, atol, RTOL_16[dn])


@pytest.mark.parametrize("shape", [(1, 16, 3, 240, 48), (2, 64, 3, 240, 48), (1, 32, 2, 312, 24), (1, 64, 2, 480, 128),
                                   (1, 128, 2, 480, 192), (1, 32, 3, 200, 130)])
def test_tcgen05_fused_regress(rsm, shape):
    """Fused correlation -> soft-argmax / argmin / argmax straight out of TMEM (bf16 features)."""
    n, c, h, w, d = shape
    rng = np.random.default_rng(41)
    l = round_to((rng.standard_normal((n, c, h, w)) * 0.5).astype(np.float32), "bf16")
    r = round_to((rng.standard_normal((n, c, h, w)) * 0.5).astype(np.float32), "bf16")
    soft, amin, amax = rsm.inner_product_regress(dev(l, "bf16"), dev(r, "bf16"), d)
    vol = oracle.inner_product_volume(l, r, d, out_dtype=np.float32)
    close(soft, oracle.soft_argmax(vol), soft_argmax_atol(d) * 4)
    assert (amax.cpu().numpy() != oracle.hard_argmax(vol)).mean() < 1e-3
    assert (amin.cpu().numpy() != oracle.hard_argmin(vol)).mean() < 1e-3
    # dyadic inputs: exact sums -> bit-exact extrema, including the x < d fill region and ties
    l = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    r = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    soft, amin, amax = rsm.inner_product_regress(dev(l, "bf16"), dev(r, "bf16"), d)
    vol = oracle.inner_product_volume(l, r, d)
    equal(amin, oracle.hard_argmin(vol))
    equal(amax, oracle.hard_argmax(vol))
    equal(rsm.inner_product_volume(dev(l, "bf16"), dev(r, "bf16"), d, out_dtype=torch.float32), vol)


@pytest.mark.parametrize("scale", [3.0, 60.0, 400.0])
def test_v4_tail_all_x4_stabiliser(rsm, scale):
    """x4 along all three axes takes the specialised staging whose softmax stabiliser is an upper BOUND
    (max of the two source rows) with an exact-maximum fallback when the bound is too loose: costs whose
    neighbouring columns differ by hundreds must still match the oracle and interpolate+softmax on the device; the
    arg-extrema too."""
    rng = np.random.default_rng(61)
    b, dc, hc, wc = 2, 48, 10, 23
    d, h, w = 4 * dc, 4 * hc, 4 * wc
    cost = (rng.standard_normal((b, dc, hc, wc)) * scale).astype(np.float32)
    cost[0, :, 3, :] = -scale * 5          # a whole coarse row far below its neighbours
    cost[0, 7, 4, ::2] = scale * 5         # isolated peaks: the bound of the rows next to them is loose
    c = dev(cost)
    pred, amin, amax = rsm.upsample_regress(c, d, h, w, argmin=True, argmax=True)
    assert torch.isfinite(pred).all()
    close(pred, oracle.v4_tail(cost, d, h, w), V4_TAIL_ATOL * max(1.0, scale / 3))
    fine = torch.nn.functional.interpolate(c.unsqueeze(1), [d, h, w], mode="trilinear").squeeze(1)
    assert (amax != fine.argmax(1)).float().mean() < 1e-3
    assert (amin != fine.argmin(1)).float().mean() < 1e-3


def test_v4_tail_bwd_wide_range(rsm):
    """The x4 head's adjoint takes its exponentials from a geometric progression unless the pixel's value range could
    flush one of them; both forms, several tiles, ragged right / bottom edges and the folded border cells are checked
    against autograd through interpolate -> softmax -> expectation on the device."""
    rng = np.random.default_rng(62)
    b, dc, hc, wc = 2, 16, 21, 45
    d, h, w = 4 * dc, 4 * hc, 4 * wc
    for scale in (2.0, 80.0):
        cost = (rng.standard_normal((b, dc, hc, wc)) * scale).astype(np.float32)
        cost[0, :, 5, :] = -scale * 4
        gout = rng.standard_normal((b, h, w)).astype(np.float32)
        c = dev(cost, grad=True)
        rsm.upsample_regress(c, d, h, w).backward(dev(gout))
        ref = dev(cost, grad=True)
        fine = torch.nn.functional.interpolate(ref.unsqueeze(1), [d, h, w], mode="trilinear").squeeze(1)
        e = (torch.softmax(fine, 1) * torch.arange(d, device="cuda", dtype=torch.float32).view(1, -1, 1, 1)).sum(1)
        e.backward(dev(gout))
        err = (c.grad - ref.grad).abs().max().item()
        assert err <= GRAD_RTOL * d * max(1.0, ref.grad.abs().max().item()), (scale, err)


@pytest.mark.parametrize("shape", [(1, 32, 3, 330, 40, 4), (2, 16, 2, 50, 24, 2), (1, 8, 2, 161, 33, 2), (1, 24, 2, 312, 48, 6),
                                   (1, 16, 3, 77, 96, 16), (2, 32, 2, 100, 20, 2), (1, 8, 2, 45, 12, 4), (1, 6, 2, 64, 16, 2)])
@pytest.mark.parametrize("dn", ["fp32", "bf16"])
def test_groupwise_bwd_row_parts(rsm, shape, dn):
    """The row-streaming adjoint kernel (1 / 2 / 4 / 8 / 16 channels per group, D % 4 == 0): ragged rows, last slab
    partly past D, several slabs, 16-bit tensors; and the shapes it leaves to the tiled kernel (D % 4 != 0, 3 channels
    per group) -- against the oracle."""
    n, c, h, w, d, ng = shape
    rng = np.random.default_rng(71)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    gout = round_to(rng.standard_normal((n, ng, h, w, d)).astype(np.float32), dn)
    gl, gr = oracle.groupwise_volume_bwd(gout, l, r, ng)
    L, R = dev(l, dn, grad=True), dev(r, dn, grad=True)
    rsm.groupwise_volume(L, R, ng, d).backward(dev(gout, dn))
    res = [(L.grad, R.grad)]
    atol = GRAD_RTOL * np.sqrt(d) * 16 if dn == "fp32" else RTOL_16[dn] * np.sqrt(d) * 4
    for a, b in res:
        close(a, gl, atol)
        close(b, gr, atol)


def test_groupwise_bwd_one_side_and_views(rsm):
    """The row-streaming adjoint with only one gradient asked for, with strided (channel-sliced, width-cropped) feature
    views (the 4-byte staging path) and with rows wider than one CTA covers (W > 1024: the tiled kernel)."""
    rng = np.random.default_rng(73)
    n, c, h, w, d, ng = 2, 16, 3, 200, 24, 4
    l = rng.standard_normal((n, c, h, w)).astype(np.float32)
    r = rng.standard_normal((n, c, h, w)).astype(np.float32)
    gout = rng.standard_normal((n, ng, h, w, d)).astype(np.float32)
    gl, gr = oracle.groupwise_volume_bwd(gout, l, r, ng)
    atol = GRAD_RTOL * np.sqrt(d) * 16
    L = dev(l, grad=True)
    rsm.groupwise_volume(L, dev(r), ng, d).backward(dev(gout))
    close(L.grad, gl, atol)
    R = dev(r, grad=True)
    rsm.groupwise_volume(dev(l), R, ng, d).backward(dev(gout))
    close(R.grad, gr, atol)
    # views: channels 4..20 of a 24-channel tensor, columns 3..203 of a 208-wide one
    big_l = torch.zeros((n, c + 8, h, w + 8), device="cuda")
    big_r = torch.zeros((n, c + 8, h, w + 8), device="cuda")
    big_l[:, 4:4 + c, :, 3:3 + w] = dev(l)
    big_r[:, 4:4 + c, :, 3:3 + w] = dev(r)
    big_l.requires_grad_(True); big_r.requires_grad_(True)
    rsm.groupwise_volume(big_l[:, 4:4 + c, :, 3:3 + w], big_r[:, 4:4 + c, :, 3:3 + w], ng, d).backward(dev(gout))
    close(big_l.grad[:, 4:4 + c, :, 3:3 + w], gl, atol)
    close(big_r.grad[:, 4:4 + c, :, 3:3 + w], gr, atol)
    assert big_l.grad[:, :4].abs().max().item() == 0 and big_l.grad[..., :3].abs().max().item() == 0
    # a row wider than the streaming kernel's CTA
    n2, c2, h2, w2, d2, g2 = 1, 8, 2, 1100, 8, 2
    l2 = rng.standard_normal((n2, c2, h2, w2)).astype(np.float32)
    r2 = rng.standard_normal((n2, c2, h2, w2)).astype(np.float32)
    go2 = rng.standard_normal((n2, g2, h2, w2, d2)).astype(np.float32)
    gl2, gr2 = oracle.groupwise_volume_bwd(go2, l2, r2, g2)
    L2, R2 = dev(l2, grad=True), dev(r2, grad=True)
    rsm.groupwise_volume(L2, R2, g2, d2).backward(dev(go2))
    close(L2.grad, gl2, atol)
    close(R2.grad, gr2, atol)


@pytest.mark.parametrize("dn", ["fp32", "bf16"])
def test_concat_bwd_disparity_chunks(rsm, dn):
    """The concatenation adjoint stages at most 64 words per pixel at a time: D = 192 (fp32: four chunks, bf16: two)
    and a D that leaves a short last chunk, ragged rows -- against the oracle."""
    rng = np.random.default_rng(74)
    for (n, c, h, w, d) in ((1, 3, 2, 230, 192), (2, 2, 3, 75, 100)):
        gout = round_to(rng.standard_normal((n, 2 * c, h, w, d)).astype(np.float32), dn)
        gl, gr = oracle.concat_volume_bwd(gout)
        L = dev(np.zeros((n, c, h, w), np.float32), dn, grad=True)
        R = dev(np.zeros((n, c, h, w), np.float32), dn, grad=True)
        rsm.concat_volume(L, R, d).backward(dev(gout, dn))
        atol = GRAD_RTOL * np.sqrt(d) * 16 if dn == "fp32" else RTOL_16[dn] * np.sqrt(d) * 4
        close(L.grad, gl, atol)
        close(R.grad, gr, atol)


@pytest.mark.parametrize("case", [(1, 4, 3, 313, 48, "fp32"), (1, 3, 2, 50, 192, "fp32"), (1, 2, 2, 1000, 48, "fp32"),
                                  (2, 6, 2, 100, 96, "bf16"), (1, 4, 2, 100, 96, "fp16"), (1, 2, 3, 64, 48, "bf16"),
                                  (1, 2, 2, 30, 48, "fp32")])
def test_concat_bwd_halves(rsm, case):
    """The concatenation adjoint by halves (pixel rows of >= 192 bytes): left half as streaming row sums, right half from
    whole rows fetched by bulk copy (fp32: one part, W not divisible into parts, D = 192, D > W) or through the right-only
    row kernel (16-bit tensors, rows too long for shared memory); 16-bit rows of 96 bytes keep the one-pass kernel.
    Dyadic gradients: the sums are exact whatever their order, so the comparison with the oracle is bit for bit."""
    n, c, h, w, d, dn = case
    rng = np.random.default_rng(75)
    gout = (rng.integers(-8, 9, (n, 2 * c, h, w, d)) / 8.0).astype(np.float32)
    gl, gr = oracle.concat_volume_bwd(gout)
    L = dev(np.zeros((n, c, h, w), np.float32), dn, grad=True)
    R = dev(np.zeros((n, c, h, w), np.float32), dn, grad=True)
    rsm.concat_volume(L, R, d).backward(dev(gout, dn))
    equal(L.grad, round_to(gl, dn))
    equal(R.grad, round_to(gr, dn))


# ------------------------------------------------------------ refinement warp
@pytest.mark.parametrize("name", names("warp_"))
def test_warp_goldens(rsm, name):
    """warp_by_flow_map against the reference's own outputs and autograd gradients."""
    g, m = load(name)
    image, flow = dev(g["image"], grad=True), dev(g["flow"], grad=True)
    out = rsm.warp_by_flow_map(image, flow)
    close(out, g["out"], 2e-5)
    out.backward(dev(g["gout"]))
    close(image.grad, g["gimage"], 5e-5)
    close(flow.grad, g["gflow"], 2e-4)


@pytest.mark.parametrize("shape", [(2, 32, 48, 156, 1), (1, 3, 96, 312, 1), (1, 5, 7, 33, 2)])
@pytest.mark.parametrize("dn", ["fp32", "bf16", "fp16"])
def test_warp_vs_oracle(rsm, shape, dn):
    """RefineNet-sized warps (v3: 32-channel features at 1/8 -> 1/4 resolution; v2: RGB) against the oracle
    and against F.grid_sample fed with the reference's grid on the device."""
    n, c, h, w, cf = shape
    rng = np.random.default_rng(81)
    image = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    flow = round_to((np.abs(rng.standard_normal((n, cf, h, w))) * 6).astype(np.float32), dn)
    out = rsm.warp_by_flow_map(dev(image, dn), dev(flow, dn))
    atol = 2e-5 if dn == "fp32" else RTOL_16[dn] * 4
    close(out, oracle.warp_by_flow_map(image, flow), atol)
    if dn == "fp32":
        im, fl = dev(image, grad=True), dev(flow, grad=True)
        gout = rng.standard_normal((n, c, h, w)).astype(np.float32)
        rsm.warp_by_flow_map(im, fl).backward(dev(gout))
        gi, gf = oracle.warp_by_flow_map_bwd(gout, image, flow)
        close(im.grad, gi, 1e-4)
        close(fl.grad, gf, 1e-4 * np.sqrt(c) * 8)
        # the reference's own op sequence on the device
        t_im, t_fl = dev(image), dev(flow)
        gy, gx = torch.meshgrid(torch.arange(h, device="cuda", dtype=torch.float32),
                                torch.arange(w, device="cuda", dtype=torch.float32), indexing="ij")
        grid_x = (gx.view(1, 1, h, w) - t_fl[:, 0].view(n, 1, h, w)).permute(0, 2, 3, 1)
        grid_y = (gy.view(1, 1, h, w) - t_fl[:, 1].view(n, 1, h, w)).permute(0, 2, 3, 1) if cf == 2 else \
            gy.view(1, h, w, 1).repeat(n, 1, 1, 1)
        grid = torch.cat((2.0 * grid_x / (w - 1.0) - 1.0, 2.0 * grid_y / (h - 1.0) - 1.0), dim=-1)
        ref = torch.nn.functional.grid_sample(t_im, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        # (ATen's CUDA kernel contracts the un-normalisation into FMAs: coordinates differ by an ulp at x ~ 150)
        torch.testing.assert_close(out, ref, atol=2e-4, rtol=0)


def test_warp_errors(rsm):
    x = torch.zeros((1, 3, 4, 8), device="cuda")
    with pytest.raises(AssertionError, match="invalid flow map dimension"):    # mobile_stereo_net_v2.py:72
        rsm.warp_by_flow_map(x, torch.zeros((1, 3, 4, 8), device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rsm.warp_by_flow_map(x.cpu(), torch.zeros((1, 1, 4, 8)))


# ------------------------------------------------------------ pre / post steps
def test_prepost_v1_goldens(rsm):
    """prepare_input / finalize_disparity against what the reference model itself computes
    (mobile_stereo_net.py:121-130, :154-159).  The goldens were produced on the CPU (true division by 255); the kernel
    follows torch's CUDA division (multiply by the fp32 reciprocal -- what the reference computes on THIS device), so
    the normalised image agrees within one ulp with the goldens and bit for bit with the oracle's device convention."""
    g, m = load("prepost_v1")
    limg = dev(g["limg"], grad=True)
    prep = rsm.prepare_input(limg, m["align"])
    ulp = 2.0 ** -23                                    # values in [-1, 1]: one ulp of the quotient doubled
    close(prep, g["prep_l"], 2 * ulp, 0)
    close(rsm.prepare_input(dev(g["rimg"]), m["align"]), g["prep_r"], 2 * ulp, 0)
    assert np.array_equal(prep.detach().cpu().numpy(), oracle.prepare_input(g["limg"], m["align"], device_div=True))
    assert np.array_equal(g["prep_l"], oracle.prepare_input(g["limg"], m["align"]))         # oracle == reference on the CPU
    prep.backward(dev(g["gprep"]))
    close(limg.grad, g["glimg"], 0, 2 * ulp)
    assert np.array_equal(limg.grad.cpu().numpy(), oracle.prepare_input_bwd(g["gprep"], g["limg"].shape[2:], device_div=True))
    padded = g["prep_l"].shape[2:]
    for k in range(m["n_out"]):
        x = dev(g[f"refined{k}"], grad=True)
        out = rsm.finalize_disparity(x, padded, (m["H"], m["W"]))
        assert torch.equal(out.cpu(), torch.from_numpy(g[f"final{k}"]))
        out.backward(dev(g[f"gfinal{k}"]))
        close(x.grad, g[f"grefined{k}"], 1e-4, 3e-4)


def test_prepost_dispnetc_goldens(rsm):
    """disparity_interpolate + crop + negate (mobile_disp_net_c.py:223-234, :408-411), six scales incl. same-size."""
    g, m = load("prepost_dispnetc")
    for k in range(m["n"]):
        x = dev(g[f"disp{k}"], grad=True)
        out = rsm.finalize_disparity(x, (m["Hp"], m["Wp"]), (m["H"], m["W"]), mode="bilinear")
        close(out, g[f"out{k}"], 1e-5, 1e-5)
        out.backward(dev(g[f"gout{k}"]))
        close(x.grad, g[f"gdisp{k}"], 1e-4, 3e-4)     # fp32 sums of up to 64 x 64 terms


@pytest.mark.parametrize("dn", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("case", [(2, 3, 384, 1248, 8), (1, 3, 375, 1242, 64), (3, 1, 17, 33, 16), (1, 3, 64, 128, 1)])
def test_prepare_vs_oracle(rsm, case, dn):
    """KITTI-sized frames (aligned: vector path; 375 x 1242 -> 384 x 1280: scalar path) and odd shapes; all dtypes
    bit-exact (16-bit tensors round after each op, as torch does)."""
    n, c, h, w, align = case
    rng = np.random.default_rng(5)
    img = round_to((rng.random((n, c, h, w)) * 255).astype(np.float32), dn)
    x = dev(img, dn)
    out = rsm.prepare_input(x, align)
    # the reference's own lines executed by torch ON THE DEVICE (CUDA division = multiply by the reciprocal): bit-exact
    ref = torch.nn.functional.pad((2.0 * (x / 255.0) - 1.0), (0, (align - w % align) % align, 0, (align - h % align) % align))
    assert out.shape == ref.shape and torch.equal(out, ref)
    if dn == "fp32":
        assert np.array_equal(out.cpu().numpy(), oracle.prepare_input(img, align, device_div=True))
        np.testing.assert_allclose(out.cpu().numpy(), oracle.prepare_input(img, align), atol=2.0 ** -22, rtol=0)   # CPU: <= 1 ulp
        gout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        xg = dev(img, grad=True)
        rsm.prepare_input(xg, align).backward(dev(gout))
        assert np.array_equal(xg.grad.cpu().numpy(), oracle.prepare_input_bwd(gout, (h, w), device_div=True))
        xt = dev(img, grad=True)
        (2.0 * (xt / 255.0) - 1.0).backward(dev(gout)[:, :, :h, :w])
        assert torch.equal(xg.grad, xt.grad)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("case", [(2, 48, 156, 384, 1248, 384, 1248), (1, 96, 320, 384, 1280, 375, 1242), (2, 5, 7, 40, 56, 37, 50),
                                   (1, 6, 9, 16, 24, 13, 22), (1, 33, 50, 20, 30, 20, 29), (1, 16, 24, 16, 24, 13, 22)])
def test_finalize_vs_oracle(rsm, case, mode):
    """Up- and down-sampling ratios, integer and fractional, cropped and not, against the oracle; the nearest map is
    bit-exact, the bilinear one and the gradients agree to fp32 rounding.  Also against F.interpolate on the device."""
    n, hs, ws, hp, wp, h, w = case
    rng = np.random.default_rng(17)
    disp = (rng.standard_normal((n, 1, hs, ws)) * 5).astype(np.float32)
    x = dev(disp, grad=True)
    out = rsm.finalize_disparity(x, (hp, wp), (h, w), mode=mode)
    ref = oracle.finalize_disparity(disp, (hp, wp), (h, w), mode)
    if mode == "nearest":
        assert np.array_equal(out.detach().cpu().numpy(), ref)
    else:
        close(out, ref, 1e-5, 1e-5)
    gout = rng.standard_normal((n, 1, h, w)).astype(np.float32)
    out.backward(dev(gout))
    close(x.grad, oracle.finalize_disparity_bwd(gout, disp.shape, (hp, wp), mode), 1e-4, 3e-4)
    t = dev(disp)
    if mode == "nearest":
        tref = -1.0 * torch.nn.functional.interpolate(t * (float(wp) / ws), (hp, wp))[:, :, :h, :w]
        assert torch.equal(out.detach(), tref)
    elif (hs, ws) != (hp, wp):
        tref = -1.0 * torch.nn.functional.interpolate(t * (float(wp) / ws), (hp, wp), mode="bilinear", align_corners=False)[:, :, :h, :w]
        torch.testing.assert_close(out.detach(), tref, atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("dn", ["bf16", "fp16"])
def test_finalize_16bit(rsm, dn):
    rng = np.random.default_rng(3)
    disp = round_to((rng.standard_normal((2, 1, 12, 20)) * 5).astype(np.float32), dn)
    x = dev(disp, dn)
    out = rsm.finalize_disparity(x, (96, 160), (90, 155))
    ref = -1.0 * torch.nn.functional.interpolate(x * (160.0 / 20), (96, 160))[:, :, :90, :155]
    assert torch.equal(out, ref)
    outb = rsm.finalize_disparity(x, (96, 160), (90, 155), mode="bilinear")
    refb = -1.0 * torch.nn.functional.interpolate(x * (160.0 / 20), (96, 160), mode="bilinear", align_corners=False)[:, :, :90, :155]
    torch.testing.assert_close(outb.float(), refb.float(), atol=RTOL_16[dn] * 64, rtol=RTOL_16[dn] * 2)


def test_prepost_errors(rsm):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rsm.prepare_input(torch.zeros((1, 3, 4, 8)), 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rsm.finalize_disparity(torch.zeros((1, 1, 4, 8)), (8, 16))
    x = torch.zeros((1, 1, 4, 8), device="cuda")
    with pytest.raises(ValueError, match="mode must be"):
        rsm.finalize_disparity(x, (8, 16), mode="bicubic")
    with pytest.raises(ValueError, match="exceeds"):
        rsm.finalize_disparity(x, (8, 16), (9, 16))
    assert rsm.prepare_input(torch.zeros((0, 3, 5, 7), device="cuda"), 8).shape == (0, 3, 8, 8)
    assert rsm.finalize_disparity(torch.zeros((0, 1, 4, 8), device="cuda"), (8, 16), (7, 15)).shape == (0, 1, 7, 15)


# ------------------------------------------------------------ loss / metrics
@pytest.mark.parametrize("name", names("loss_"))
def test_loss_goldens(rsm, name):
    """SequenceLoss / get_flow_map_metrics mirrors against the reference's own loss value, gradients and metrics."""
    g, m = load(name)
    preds = [dev(g[f"pred{k}"], grad=True) for k in range(m["n_preds"])]
    gt, valid = dev(g["gt"]), dev(g["valid"])
    loss = rsm.SequenceLoss(loss_gamma=m["gamma"], max_flow_magnitude=m["max_flow"])(preds, gt, valid)
    close(loss, g["loss"], 0, 2e-6)
    loss.backward()
    for k, p in enumerate(preds):
        close(p.grad, g[f"gpred{k}"], 1e-8, 1e-5)
    got = rsm.get_flow_map_metrics(gt, preds[-1].detach(), valid)
    assert list(got) == list(m["metrics"])
    for key, ref in m["metrics"].items():
        np.testing.assert_allclose(got[key], ref, atol=1e-7, rtol=2e-6)


@pytest.mark.parametrize("case", [(8, 384, 1248, [(96, 312), (192, 624), (384, 1248)]), (2, 375, 1242, [(375, 1242)] * 3),
                                   (1, 33, 77, [(5, 9), (33, 77)])])
def test_loss_vs_oracle(rsm, case):
    """Training-batch sizes (cfg3 frames; DispNetC-style full-size predictions; odd ratios) against the oracle,
    and against the reference's op sequence on the device."""
    n, h, w, sizes = case
    rng = np.random.default_rng(11)
    gt = (rng.standard_normal((n, 1, h, w)) * 40).astype(np.float32)
    valid = (rng.random((n, h, w)) > 0.2).astype(np.float32)
    preds = [(rng.standard_normal((n, 1) + s) * 10).astype(np.float32) for s in sizes]
    tp = [dev(p, grad=True) for p in preds]
    tgt, tv = dev(gt), dev(valid)
    loss = rsm.SequenceLoss(0.9, 60.0)(tp, tgt, tv)
    close(loss, oracle.sequence_loss(preds, gt, valid, 0.9, 60.0), 0, 2e-6)
    loss.backward()
    for p, gp in zip(tp, oracle.sequence_loss_bwd(preds, gt, valid, 0.9, 60.0)):
        close(p.grad, gp, 1e-9, 1e-5)
    # the reference's op sequence with torch on the device
    Fn = torch.nn.functional
    mask = ((tv >= 0.5) & (torch.sum(tgt ** 2, dim=1).sqrt() < 60.0)).unsqueeze(1)
    ref = 0.0
    for i, p in enumerate(tp):
        q = p.detach()
        if q.shape != tgt.shape:
            q = Fn.interpolate(q * (float(w) / q.shape[-1]), (h, w))
        el = Fn.smooth_l1_loss(tgt, q, reduction="none", beta=1.0) if i == len(tp) - 1 else Fn.l1_loss(tgt, q, reduction="none")
        ref = ref + 0.9 ** (len(tp) - 1 - i) * el[mask].mean()
    torch.testing.assert_close(loss.detach(), ref, rtol=2e-5, atol=0)
    got = rsm.get_flow_map_metrics(tgt, tp[-1].detach(), tv)
    want = oracle.flow_map_metrics(gt, preds[-1], valid)
    for key in want:
        np.testing.assert_allclose(got[key], want[key], atol=1e-7, rtol=2e-6)


@pytest.mark.parametrize("c", [1, 2, 3])
def test_flow_metrics_channels(rsm, c):
    """get_flow_map_metrics sums the squared error over the channel axis (loss.py:10): optical-flow shaped inputs."""
    rng = np.random.default_rng(23 + c)
    gt = (rng.standard_normal((2, c, 19, 31)) * 3).astype(np.float32)
    pred = gt + rng.standard_normal(gt.shape).astype(np.float32) * 2
    valid = (rng.random((2, 19, 31)) > 0.4).astype(np.float32)
    got = rsm.get_flow_map_metrics(dev(gt), dev(pred), dev(valid))
    want = oracle.flow_map_metrics(gt, pred, valid)
    for key in want:
        np.testing.assert_allclose(got[key], want[key], atol=1e-7, rtol=3e-6, err_msg=key)
    vec = rsm.flow_map_metrics(dev(gt), dev(pred), dev(valid))
    assert vec.dtype == torch.float64 and vec.shape == (8,) and float(vec[7]) == float(valid.sum())


def test_loss_error_behaviour(rsm):
    gt = torch.zeros((1, 1, 4, 8), device="cuda")
    valid = torch.ones((1, 4, 8), device="cuda")
    loss = rsm.SequenceLoss()
    with pytest.raises(AssertionError, match="empty flow predictions"):          # loss.py:53
        loss([], gt, valid)
    with pytest.raises(AssertionError):                                           # loss.py:59
        loss([gt], gt, torch.ones((1, 4, 7), device="cuda"))
    bad = gt.clone()
    bad[0, 0, 1, 2] = float("nan")
    with pytest.raises(AssertionError, match="non-finite"):                       # loss.py:66-67
        loss([bad], gt, valid)
    assert torch.isnan(rsm.SequenceLoss(check_finite=False)([bad], gt, valid))
    assert torch.isnan(loss([gt], gt, torch.zeros_like(valid)))                   # empty mask: mean of nothing
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss([gt.cpu()], gt.cpu(), valid.cpu())
    with pytest.raises(NotImplementedError):
        rsm.build_loss_function({"type": "Other", "parameters": {}})
    assert isinstance(rsm.build_loss_function({"type": "SequenceLoss", "parameters": {"loss_gamma": 0.8}}), rsm.SequenceLoss)


def test_empty_and_degenerate_inputs(rsm):
    """Empty batch / zero-width inputs and D = 0 go through every op without touching memory (SURVEY 8c:
    'empty and ragged inputs'); shapes follow the reference's."""
    z = torch.zeros((0, 8, 4, 16), device="cuda")
    assert rsm.concat_volume(z, z, 3).shape == (0, 16, 4, 16, 3)
    assert rsm.interweave(z, z).shape == (0, 16, 4, 16)
    assert rsm.inner_product_volume(z, z, 3).shape == (0, 3, 4, 16)
    assert rsm.groupwise_volume(z, z, 4, 3).shape == (0, 4, 4, 16, 3)
    assert rsm.difference_volume(z, z, 3).shape == (0, 8, 3, 4, 16)
    assert rsm.soft_argmax(torch.zeros((0, 5, 4, 16), device="cuda")).shape == (0, 4, 16)
    assert rsm.upsample_regress(torch.zeros((0, 3, 2, 4), device="cuda"), 12, 8, 16).shape == (0, 8, 16)
    assert rsm.warp_by_flow_map(z, torch.zeros((0, 1, 4, 16), device="cuda")).shape == (0, 8, 4, 16)
    x = torch.randn((1, 8, 4, 16), device="cuda", requires_grad=True)
    v = rsm.concat_volume(x, x, 0)                      # no disparities: empty last axis, gradient is zero
    assert v.shape == (1, 16, 4, 16, 0)
    v.sum().backward()
    assert torch.equal(x.grad, torch.zeros_like(x))
    w0 = torch.zeros((2, 8, 4, 0), device="cuda")       # zero-width feature maps
    assert rsm.inner_product_volume(w0, w0, 3).shape == (2, 3, 4, 0)
    assert rsm.difference_volume(w0, w0, 3).shape == (2, 8, 3, 4, 0)


@pytest.mark.parametrize("shape", [(1, 64, 3, 240, 48), (2, 40, 2, 136, 70), (1, 32, 2, 128, 24)])
@pytest.mark.parametrize("dn", ["bf16", "fp16"])
@pytest.mark.parametrize("mean", [False, True])
def test_inner_bwd_16bit_big_tiles(rsm, shape, dn, mean):
    """16-bit tensors through the 8x8-tile inner-product adjoint (rows on 16-byte boundaries: W % 8 == 0),
    against the fp32 oracle on the same rounded inputs."""
    n, c, h, w, d = shape
    rng = np.random.default_rng(91)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    gout = round_to(rng.standard_normal((n, d, h, w)).astype(np.float32), dn)
    gl, gr = oracle.inner_product_volume_bwd(gout, l, r, mean=mean)
    atol = RTOL_16[dn] * np.sqrt(d) * 4 / (c if mean else 1)
    L, R = dev(l, dn, grad=True), dev(r, dn, grad=True)
    rsm.inner_product_volume(L, R, d, mean=mean).backward(dev(gout, dn))
    close(L.grad, gl, atol, RTOL_16[dn])
    close(R.grad, gr, atol, RTOL_16[dn])


@pytest.mark.parametrize("shape", [(3, 64, 130, 132, 48), (2, 48, 40, 260, 96), (1, 40, 5, 132, 70), (2, 32, 7, 128, 20),
                                   (1, 16, 2, 516, 200)])
@pytest.mark.parametrize("mean", [False, True])
def test_inner_bwd_fp32_stream(rsm, shape, mean):
    """fp32 inner-product adjoint, persistent two-stage kernel (inner_bwd_stream_kernel): more tiles than CTAs (every
    CTA walks several items), one / two / four disparity chunks, a ragged last channel block, rows that end inside a
    tile -- against the oracle, and bit-exact on dyadic inputs (fp32 sums are then order-independent)."""
    n, c, h, w, d = shape
    rng = np.random.default_rng(17)
    for dyadic in (False, True):
        if dyadic:
            l, r, gout = ((rng.integers(-8, 9, s) / 8.0).astype(np.float32) for s in ((n, c, h, w), (n, c, h, w), (n, d, h, w)))
        else:
            l, r, gout = (rng.standard_normal(s).astype(np.float32) for s in ((n, c, h, w), (n, c, h, w), (n, d, h, w)))
        if h * w * n > 20000:      # the numpy oracle walks disparities: keep it to a few rows of the big cases
            rows = slice(0, h, max(1, h // 3))
        else:
            rows = slice(None)
        gl, gr = oracle.inner_product_volume_bwd(gout[:, :, rows], l[:, :, rows], r[:, :, rows], mean=mean)
        L, R = dev(l, grad=True), dev(r, grad=True)
        rsm.inner_product_volume(L, R, d, mean=mean).backward(dev(gout))
        if dyadic and (not mean or c & (c - 1) == 0):
            equal(L.grad[:, :, rows], gl)
            equal(R.grad[:, :, rows], gr)
        else:
            atol = GRAD_RTOL * np.sqrt(d) * 16 / (c if mean else 1)
            close(L.grad[:, :, rows], gl, atol)
            close(R.grad[:, :, rows], gr, atol)
