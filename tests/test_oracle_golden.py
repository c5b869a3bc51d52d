"""CPU suite: pin the numpy oracle against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

import oracle
from golden_io import load, names, round_to
from tolerances import (GRAD_RTOL, RTOL_16, V4_TAIL_ATOL, corr_atol_fp32, soft_argmax_atol)


def _close(a, b, atol, rtol=0.0):
    np.testing.assert_allclose(a, b, atol=atol, rtol=rtol)


@pytest.mark.parametrize("name", names("vol_"))
def test_volume_ops(name):
    g, m = load(name)
    l, r, d, ng, c = g["left"], g["right"], m["D"], m["G"], m["C"]
    dn, exact = m["dtype"], m["kind"] == "dyadic"
    # ---- pure data movement: always bit exact
    np.testing.assert_array_equal(oracle.concat_volume(l, r, d), g["concat.out"])
    np.testing.assert_array_equal(oracle.interweave(l, r), g["interweave.out"])
    np.testing.assert_array_equal(oracle.interweave(l, r), g["interweave_v4.out"])
    np.testing.assert_array_equal(round_to(oracle.difference_volume(l, r, d), dn), g["difference.out"])
    gl, gr = oracle.interweave_bwd(g["interweave.gout"])
    np.testing.assert_array_equal(gl, g["interweave.gleft"])
    np.testing.assert_array_equal(gr, g["interweave.gright"])
    # ---- reductions
    lmax, rmax = np.abs(l).max(), np.abs(r).max()
    atol = 0.0 if (exact and dn == "fp32") else corr_atol_fp32(c, lmax, rmax)
    rtol = 0.0 if dn == "fp32" else RTOL_16[dn] * 2
    if dn != "fp32":
        atol = RTOL_16[dn] * np.sqrt(c) * lmax * rmax
    _close(oracle.inner_product_volume(l, r, d), g["inner.out"], atol, rtol)
    _close(oracle.inner_product_volume(l, r, d, mean=True), g["corr_mean.out"], atol, rtol)
    # the reference's groupwise output is always fp32 but its products are
    # rounded to the input dtype first
    _close(oracle.groupwise_volume(l, r, ng, d), g["groupwise.out"], atol, rtol)
    # ---- gradients (fp32 goldens only: 16-bit autograd accumulates in 16 bit)
    if dn != "fp32":
        return
    D = d
    gatol = 0.0 if exact else GRAD_RTOL * max(1.0, np.sqrt(D))
    for op, (gl, gr) in {
        "concat": oracle.concat_volume_bwd(g["concat.gout"]),
        "difference": oracle.difference_volume_bwd(g["difference.gout"]),
    }.items():
        _close(gl, g[f"{op}.gleft"], GRAD_RTOL * np.sqrt(D))
        _close(gr, g[f"{op}.gright"], GRAD_RTOL * np.sqrt(D))
    gs = GRAD_RTOL * np.sqrt(D) * max(lmax, rmax) * 4
    gl, gr = oracle.inner_product_volume_bwd(g["inner.gout"], l, r)
    _close(gl, g["inner.gleft"], gs), _close(gr, g["inner.gright"], gs)
    gl, gr = oracle.inner_product_volume_bwd(g["corr_mean.gout"], l, r, mean=True)
    _close(gl, g["corr_mean.gleft"], gs), _close(gr, g["corr_mean.gright"], gs)
    gl, gr = oracle.groupwise_volume_bwd(g["groupwise.gout"], l, r, ng)
    _close(gl, g["groupwise.gleft"], gs), _close(gr, g["groupwise.gright"], gs)


def test_noncontiguous_slices():
    g, m = load("noncontig_interweave")
    i = m["i"]
    a, b = g["featL"][:, :, :, i:], g["featR"][:, :, :, :-i]
    np.testing.assert_array_equal(oracle.interweave(a, b), g["out"])
    np.testing.assert_array_equal(oracle.concat_volume(a, b, 5), g["concat"])
    _close(oracle.inner_product_volume(a, b, 5), g["inner"], 1e-5)


@pytest.mark.parametrize("name", names("regress_"))
def test_regression(name):
    g, m = load(name)
    cost = g["cost"]
    np.testing.assert_array_equal(oracle.hard_argmin(cost), g["argmin"])
    np.testing.assert_array_equal(oracle.hard_argmax(cost), g["argmax"])
    if "e" not in g:
        return
    dn = m["dtype"]
    atol = soft_argmax_atol(m["D"]) if dn == "fp32" else RTOL_16[dn] * m["D"]
    _close(oracle.soft_argmax(cost), g["e"], atol)
    _close(oracle.soft_argmax(cost, keepdim=True), g["e_keepdim"], atol)
    if dn == "fp32":
        _close(oracle.soft_argmax_bwd(g["gout"], cost), g["gcost"], GRAD_RTOL * m["D"])


@pytest.mark.parametrize("name", names("tail_"))
def test_v4_tail(name):
    g, m = load(name)
    cost = g["cost"]
    fine = oracle.trilinear_upsample(cost, m["D"], m["H"], m["W"])
    _close(fine, g["fine"], 2e-5)
    _close(oracle.v4_tail(cost, m["D"], m["H"], m["W"]), g["pred"], V4_TAIL_ATOL)
    _close(oracle.v4_tail_bwd(g["gout"], cost, m["D"], m["H"], m["W"]), g["gcost"],
           GRAD_RTOL * m["D"])
    np.testing.assert_array_equal(oracle.hard_argmax(g["fine"]), g["argmax"])
    np.testing.assert_array_equal(oracle.hard_argmin(g["fine"]), g["argmin"])


def test_callsite_v1():
    """model/mobile_stereo_net.py:140 (difference volume) and :144-147 (regression)."""
    g, m = load("callsite_v1")
    np.testing.assert_array_equal(oracle.difference_volume(g["lf"], g["rf"], m["max_disp"]), g["volume"])
    _close(oracle.soft_argmax(g["filtered"], keepdim=True), g["regressed"], soft_argmax_atol(m["max_disp"]))


def test_callsite_dispnetc():
    """model/mobile_disp_net_c.py:365-367 (mean correlation at 1/4 res)."""
    g, m = load("callsite_dispnetc")
    c = g["lf"].shape[1]
    atol = corr_atol_fp32(c, np.abs(g["lf"]).max(), np.abs(g["rf"]).max())
    _close(oracle.inner_product_volume(g["lf"], g["rf"], m["max_disp"], mean=True), g["volume"], atol)


def test_callsite_v4():
    """model/mobile_stereo_net_v4.py:446/:453 (interweave) and :511-520 (tail, negated)."""
    g, m = load("callsite_v4")
    for k in range(m["n_iw"]):
        np.testing.assert_array_equal(oracle.interweave(g[f"iw{k}.a"], g[f"iw{k}.b"]), g[f"iw{k}.out"])
    pred = oracle.v4_tail(g["cost3"], m["maxdisp"], m["H"], m["W"])
    _close(-pred[:, None], g["final"], V4_TAIL_ATOL)


def test_groupwise_assert():
    l = np.zeros((1, 6, 2, 3), np.float32)
    with pytest.raises(AssertionError):
        oracle.groupwise_volume(l, l, 4, 2)


@pytest.mark.parametrize("name", names("warp_"))
def test_warp_goldens(name):
    """warp_by_flow_map, model/mobile_stereo_net_v2.py:59-96 (= v3 :60-97): output and autograd gradients of
    the reference itself (1- and 2-channel flows, samples leaving the image on both sides)."""
    g, m = load(name)
    _close(oracle.warp_by_flow_map(g["image"], g["flow"]), g["out"], 2e-5)
    gi, gf = oracle.warp_by_flow_map_bwd(g["gout"], g["image"], g["flow"])
    _close(gi, g["gimage"], 2e-5)
    _close(gf, g["gflow"], 1e-4)


def test_warp_assert():
    with pytest.raises(AssertionError, match="invalid flow map dimension"):
        oracle.warp_by_flow_map(np.zeros((1, 2, 3, 4), np.float32), np.zeros((1, 3, 3, 4), np.float32))


# ------------------------------------------------------------ pre / post steps
def test_prepost_v1_goldens():
    """The reference model's own prepared images (input of feature_extractor) and the mapping from each RefineNet
    output to the returned map, mobile_stereo_net.py:121-130 / :154-159, with autograd gradients."""
    g, m = load("prepost_v1")
    for img, prep in ((g["limg"], g["prep_l"]), (g["rimg"], g["prep_r"])):
        out = oracle.prepare_input(img, m["align"])
        assert out.shape == prep.shape and np.array_equal(out, prep)          # bit-exact: same fp32 op sequence
    assert np.array_equal(oracle.prepare_input_bwd(g["gprep"], (m["H"], m["W"])), g["glimg"])
    padded = g["prep_l"].shape[2:]
    for k in range(m["n_out"]):
        out = oracle.finalize_disparity(g[f"refined{k}"], padded, (m["H"], m["W"]), "nearest")
        assert np.array_equal(out, g[f"final{k}"])                              # a gather: bit-exact
        _close(oracle.finalize_disparity_bwd(g[f"gfinal{k}"], g[f"refined{k}"].shape, padded, "nearest"),
               g[f"grefined{k}"], 1e-4, 3e-4)


def test_prepost_dispnetc_goldens():
    """disparity_interpolate + crop + negate, mobile_disp_net_c.py:223-234 + :408-411, six scales."""
    g, m = load("prepost_dispnetc")
    for k in range(m["n"]):
        out = oracle.finalize_disparity(g[f"disp{k}"], (m["Hp"], m["Wp"]), (m["H"], m["W"]), "bilinear")
        _close(out, g[f"out{k}"], 1e-5, 1e-5)
        _close(oracle.finalize_disparity_bwd(g[f"gout{k}"], g[f"disp{k}"].shape, (m["Hp"], m["Wp"]), "bilinear"),
               g[f"gdisp{k}"], 1e-4, 3e-4)     # fp32 sums of up to 64 x 64 terms


# ------------------------------------------------------------ loss / metrics
@pytest.mark.parametrize("name", names("loss_"))
def test_loss_goldens(name):
    """SequenceLoss.forward / get_flow_map_metrics, loss/loss.py:6-81, against the reference's own values and
    autograd gradients (pyramid of predictions, fractional resize ratio, max-flow and validity masks)."""
    g, m = load(name)
    preds = [g[f"pred{k}"] for k in range(m["n_preds"])]
    _close(oracle.sequence_loss(preds, g["gt"], g["valid"], m["gamma"], m["max_flow"]), g["loss"], 0, 1e-6)
    for k, gp in enumerate(oracle.sequence_loss_bwd(preds, g["gt"], g["valid"], m["gamma"], m["max_flow"])):
        _close(gp, g[f"gpred{k}"], 1e-8, 1e-5)
    got = oracle.flow_map_metrics(g["gt"], preds[-1], g["valid"])
    for key, ref in m["metrics"].items():
        _close(got[key], ref, 1e-7, 1e-6)
