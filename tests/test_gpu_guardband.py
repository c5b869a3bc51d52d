"""GPU suite: out-of-bounds write detection without compute-sanitizer (closed on this pool).

Every forward/backward entry point of the C ABI is called DIRECTLY (ctypes, raw pointers) with its
output placed in the middle of a larger buffer pre-filled with a sentinel; after the call the guard
bands on both sides must be untouched and the payload must be fully overwritten (no sentinel
left).  Shapes are ragged on purpose (nothing divisible by the vector widths, D > W, tile tails)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SENT = -61440.0          # exactly representable in fp32 / fp16 / bf16 and > 10 sigma away from anything an op
                         # can legitimately produce here (a sentinel inside the output range makes 'never written'
                         # fire at random: -12345.5 rounds to -12352 in bf16, which sum_d d*cost[d] hits ~3 % of the time)
GUARD = 1024             # elements on each side


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(20240917)     # deterministic inputs: a guard-band failure must be reproducible


@pytest.fixture(scope="module")
def L():
    from realtime_stereo_matcher_b200 import _lib
    _lib.load()
    return _lib


class Guarded:
    def __init__(self, numel, dtype):
        self.buf = torch.full((numel + 2 * GUARD,), SENT, dtype=dtype, device="cuda")
        self.numel = numel
        self.payload = self.buf[GUARD:GUARD + numel]

    def ptr(self):
        return self.payload.data_ptr()

    def check(self, what, expect_full=True):
        torch.cuda.synchronize()
        lo, hi = self.buf[:GUARD], self.buf[GUARD + self.numel:]
        assert bool((lo == SENT).all()) and bool((hi == SENT).all()), f"{what}: wrote outside its output"
        if expect_full and self.numel:
            left = int((self.payload == SENT).sum())
            assert left == 0, f"{what}: {left} output elements never written"


SHAPES = [(2, 12, 5, 67, 19, 3), (1, 16, 3, 130, 70, 4), (1, 32, 2, 240, 48, 8), (3, 5, 4, 9, 13, 5), (1, 8, 1, 1, 3, 2),
          # concat adjoint by halves: 16-bit rows of >= 192 bytes (left streaming kernel + right-only row kernel), fp32 rows
          # fetched by bulk copy whole (W not divisible into parts) -- (1, 32, 2, 240, 48, 8) above takes it in two parts
          (1, 8, 2, 100, 96, 4), (1, 4, 2, 313, 48, 2)]
DTYPES = [(torch.float32, 0), (torch.float16, 1), (torch.bfloat16, 2)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dt", DTYPES)
def test_volume_entry_points_stay_in_bounds(L, shape, dt):
    n, c, h, w, d, g = shape
    tdt, code = dt
    lib = L.load()
    gen = torch.Generator(device="cuda").manual_seed(3)
    left = torch.randn((n, c, h, w), device="cuda", generator=gen).to(tdt)
    right = torch.randn((n, c, h, w), device="cuda", generator=gen).to(tdt)
    fl, fr = L.feat(left), L.feat(right)
    st = L.stream_ptr(0)

    out = Guarded(n * 2 * c * h * w * d, tdt)
    L.check(lib.rsm_concat_fwd(fl, fr, out.ptr(), n, c, h, w, d, code, 0, st), "concat")
    out.check("rsm_concat_fwd")
    gl, gr = Guarded(n * c * h * w, tdt), Guarded(n * c * h * w, tdt)
    L.check(lib.rsm_concat_bwd(out.ptr(), gl.ptr(), gr.ptr(), n, c, h, w, d, code, 0, st), "concat_bwd")
    gl.check("rsm_concat_bwd gl"), gr.check("rsm_concat_bwd gr")

    out = Guarded(n * 2 * c * h * w, tdt)
    L.check(lib.rsm_interweave_fwd(fl, fr, out.ptr(), n, c, h, w, code, 0, st), "interweave")
    out.check("rsm_interweave_fwd")
    gl, gr = Guarded(n * c * h * w, tdt), Guarded(n * c * h * w, tdt)
    L.check(lib.rsm_interweave_bwd(out.ptr(), gl.ptr(), gr.ptr(), n, c, h, w, code, 0, st), "interweave_bwd")
    gl.check("rsm_interweave_bwd gl"), gr.check("rsm_interweave_bwd gr")

    out = Guarded(n * c * d * h * w, tdt)
    L.check(lib.rsm_difference_fwd(fl, fr, out.ptr(), n, c, h, w, d, 1.0, code, 0, st), "difference")
    out.check("rsm_difference_fwd")
    gl, gr = Guarded(n * c * h * w, tdt), Guarded(n * c * h * w, tdt)
    L.check(lib.rsm_difference_bwd(out.ptr(), gl.ptr(), gr.ptr(), n, c, h, w, d, code, 0, st), "difference_bwd")
    gl.check("rsm_difference_bwd gl"), gr.check("rsm_difference_bwd gr")

    for odt, ocode in ((tdt, code), (torch.float32, 0)):
        out = Guarded(n * d * h * w, odt)
        L.check(lib.rsm_inner_fwd(fl, fr, out.ptr(), n, c, h, w, d, 1, code, ocode, 0, st), "inner")
        out.check("rsm_inner_fwd")
        gl, gr = Guarded(n * c * h * w, tdt), Guarded(n * c * h * w, tdt)
        L.check(lib.rsm_inner_bwd(out.ptr(), fl, fr, gl.ptr(), gr.ptr(), n, c, h, w, d, 1, code, ocode, 0, st), "inner_bwd")
        gl.check("rsm_inner_bwd gl"), gr.check("rsm_inner_bwd gr")
        if c % g == 0:
            out = Guarded(n * g * h * w * d, odt)
            L.check(lib.rsm_groupwise_fwd(fl, fr, out.ptr(), n, c, h, w, d, g, code, ocode, 0, st), "groupwise")
            out.check("rsm_groupwise_fwd")
            gl, gr = Guarded(n * c * h * w, tdt), Guarded(n * c * h * w, tdt)
            L.check(lib.rsm_groupwise_bwd(out.ptr(), fl, fr, gl.ptr(), gr.ptr(), n, c, h, w, d, g, code, ocode, 0, st),
                    "groupwise_bwd")
            gl.check("rsm_groupwise_bwd gl"), gr.check("rsm_groupwise_bwd gr")

    if d <= 512:
        so, mi, ma = Guarded(n * h * w, torch.float32), Guarded(n * h * w, torch.float32), Guarded(n * h * w, torch.float32)
        # argmin/argmax are int64: place them in float64-sized guarded buffers via separate tensors
        mi64 = torch.full((n * h * w + 2 * GUARD,), -7, dtype=torch.int64, device="cuda")
        ma64 = torch.full((n * h * w + 2 * GUARD,), -7, dtype=torch.int64, device="cuda")
        ro = L.RsmRegressOut(so.ptr(), mi64[GUARD:].data_ptr(), ma64[GUARD:].data_ptr(), None)
        L.check(lib.rsm_inner_regress_fwd(fl, fr, n, c, h, w, d, 0, code, ro, 0, st), "inner_regress")
        so.check("rsm_inner_regress_fwd soft")
        torch.cuda.synchronize()
        for t in (mi64, ma64):
            assert bool((t[:GUARD] == -7).all()) and bool((t[GUARD + n * h * w:] == -7).all())
            assert bool((t[GUARD:GUARD + n * h * w] >= 0).all())


@pytest.mark.parametrize("shape", [(2, 19, 5, 13), (1, 192, 7, 33), (1, 1, 3, 5)])
@pytest.mark.parametrize("dt", DTYPES)
def test_regress_entry_points_stay_in_bounds(L, shape, dt):
    n, d, h, w = shape
    tdt, code = dt
    lib = L.load()
    cost = (torch.randn((n, d, h, w), device="cuda") * 3).to(tdt)
    st = L.stream_ptr(0)
    so, ls = Guarded(n * h * w, tdt), Guarded(n * h * w, torch.float32)
    mi64 = torch.full((n * h * w + 2 * GUARD,), -7, dtype=torch.int64, device="cuda")
    e32 = Guarded(n * h * w, torch.float32)
    ro = L.RsmRegressOut(so.ptr(), mi64[GUARD:].data_ptr(), None, ls.ptr(), e32.ptr())
    L.check(lib.rsm_regress_fwd(cost.data_ptr(), n, d, h, w, code, ro, 0, st), "regress")
    so.check("rsm_regress_fwd soft"), ls.check("rsm_regress_fwd lse"), e32.check("rsm_regress_fwd expect")
    assert torch.equal(e32.payload.to(tdt), so.payload)       # soft is the fp32 expectation rounded once
    assert bool((mi64[:GUARD] == -7).all()) and bool((mi64[GUARD + n * h * w:] == -7).all())
    gout = torch.randn((n, h, w), device="cuda").to(tdt)
    gc = Guarded(n * d * h * w, tdt)
    L.check(lib.rsm_regress_bwd(gout.data_ptr(), cost.data_ptr(), e32.ptr(), ls.ptr(), gc.ptr(), n, d, h, w, code, 0, st),
            "regress_bwd")
    gc.check("rsm_regress_bwd")
    ex = Guarded(n * h * w, tdt)
    L.check(lib.rsm_expect_fwd(cost.data_ptr(), ex.ptr(), n, d, h, w, code, 0, st), "expect")
    ex.check("rsm_expect_fwd")
    gp = Guarded(n * d * h * w, tdt)
    L.check(lib.rsm_expect_bwd(gout.data_ptr(), gp.ptr(), n, d, h, w, code, 0, st), "expect_bwd")
    gp.check("rsm_expect_bwd")


@pytest.mark.parametrize("geom", [(2, 12, 5, 7, 48, 20, 28), (1, 5, 3, 4, 13, 7, 10), (1, 9, 8, 8, 9, 4, 4), (1, 48, 3, 9, 192, 12, 36)])
def test_v4_head_entry_points_stay_in_bounds(L, geom):
    b, dc, hc, wc, d, h, w = geom
    lib = L.load()
    cost = torch.randn((b, dc, hc, wc), device="cuda") * 3
    st = L.stream_ptr(0)
    so, ls = Guarded(b * h * w, torch.float32), Guarded(b * h * w, torch.float32)
    ro = L.RsmRegressOut(so.ptr(), None, None, ls.ptr(), so.ptr())      # fp32 cost: soft and expect may alias
    L.check(lib.rsm_upsample_regress_fwd(cost.data_ptr(), b, dc, hc, wc, d, h, w, 0, ro, 0, st), "tail")
    so.check("rsm_upsample_regress_fwd soft"), ls.check("rsm_upsample_regress_fwd lse")
    gout = torch.randn((b, h, w), device="cuda")
    nbytes = lib.rsm_upsample_regress_bwd_workspace(b, dc, h, w)
    work, gc = Guarded(nbytes // 4, torch.float32), Guarded(b * dc * hc * wc, torch.float32)
    L.check(lib.rsm_upsample_regress_bwd(gout.data_ptr(), cost.data_ptr(), so.ptr(), ls.ptr(), gc.ptr(), work.ptr(),
                                         b, dc, hc, wc, d, h, w, 0, 0, st), "tail_bwd")
    work.check("rsm_upsample_regress_bwd workspace"), gc.check("rsm_upsample_regress_bwd gcost")


@pytest.mark.parametrize("shape", [(2, 5, 7, 33, 1), (1, 3, 4, 9, 2), (1, 1, 2, 2, 1)])
@pytest.mark.parametrize("dt", DTYPES)
def test_warp_entry_points_stay_in_bounds(L, shape, dt):
    n, c, h, w, cf = shape
    tdt, code = dt
    lib = L.load()
    st = L.stream_ptr(0)
    image = torch.randn((n, c, h, w), device="cuda").to(tdt)
    flow = (torch.randn((n, cf, h, w), device="cuda") * 5).to(tdt)       # many samples leave the image
    out = Guarded(n * c * h * w, tdt)
    L.check(lib.rsm_warp_fwd(image.data_ptr(), flow.data_ptr(), out.ptr(), n, c, h, w, cf, code, 0, st), "warp")
    out.check("rsm_warp_fwd")
    gimage, gflow = Guarded(n * c * h * w, torch.float32), Guarded(n * cf * h * w, tdt)
    L.check(lib.rsm_warp_bwd(out.ptr(), image.data_ptr(), flow.data_ptr(), gimage.ptr(), gflow.ptr(), n, c, h, w, cf,
                             code, 0, st), "warp_bwd")
    gimage.check("rsm_warp_bwd gimage"), gflow.check("rsm_warp_bwd gflow")


@pytest.mark.parametrize("case", [(3, 13, 22, 16, 24), (2, 8, 16, 8, 16), (1, 5, 7, 8, 8)])
@pytest.mark.parametrize("dt", DTYPES)
def test_prepare_entry_points_stay_in_bounds(L, case, dt):
    planes, h, w, hp, wp = case
    tdt, code = dt
    lib = L.load()
    st = L.stream_ptr(0)
    img = (torch.rand((planes, h, w), device="cuda") * 255).to(tdt)
    out = Guarded(planes * hp * wp, tdt)
    L.check(lib.rsm_prepare_fwd(img.data_ptr(), out.ptr(), planes, h, w, hp, wp, code, 0, st), "prepare")
    out.check("rsm_prepare_fwd")
    gimg = Guarded(planes * h * w, tdt)
    L.check(lib.rsm_prepare_bwd(out.ptr(), gimg.ptr(), planes, h, w, hp, wp, code, 0, st), "prepare_bwd")
    gimg.check("rsm_prepare_bwd")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", [(2, 5, 7, 40, 56, 37, 50), (1, 33, 50, 20, 30, 20, 29), (3, 1, 1, 9, 5, 9, 5), (1, 6, 9, 6, 9, 5, 9)])
@pytest.mark.parametrize("dt", DTYPES)
def test_finalize_entry_points_stay_in_bounds(L, case, mode, dt):
    planes, hs, ws, hp, wp, h, w = case
    tdt, code = dt
    lib = L.load()
    st = L.stream_ptr(0)
    disp = torch.randn((planes, hs, ws), device="cuda").to(tdt)
    out = Guarded(planes * h * w, tdt)
    L.check(lib.rsm_finalize_fwd(disp.data_ptr(), out.ptr(), planes, hs, ws, hp, wp, h, w, wp / ws, mode, code, 0, st), "finalize")
    out.check("rsm_finalize_fwd")
    gdisp = Guarded(planes * hs * ws, tdt)
    L.check(lib.rsm_finalize_bwd(out.ptr(), gdisp.ptr(), planes, hs, ws, hp, wp, h, w, wp / ws, mode, code, 0, st), "finalize_bwd")
    gdisp.check("rsm_finalize_bwd")


@pytest.mark.parametrize("case", [(2, 6, 10, 24, 40), (1, 7, 11, 20, 30), (3, 9, 13, 9, 13)])
@pytest.mark.parametrize("dt", DTYPES)
def test_loss_entry_points_stay_in_bounds(L, case, dt):
    n, hs, ws, h, w = case
    tdt, code = dt
    lib = L.load()
    st = L.stream_ptr(0)
    pred = torch.randn((n, 1, hs, ws), device="cuda").to(tdt)
    gt = (torch.randn((n, 1, h, w), device="cuda") * 3).to(tdt)
    valid = (torch.rand((n, h, w), device="cuda") > 0.3).float()
    work = torch.empty((L.RSM_REDUCE_WS_DOUBLES,), dtype=torch.float64, device="cuda")
    for kind in (0, 1):
        result = Guarded(4, torch.float64)
        L.check(lib.rsm_seqloss_fwd(pred.data_ptr(), gt.data_ptr(), valid.data_ptr(), work.data_ptr(), result.ptr(), n, hs, ws,
                                    h, w, 700.0, kind, code, 0, st), "seqloss")
        result.check("rsm_seqloss_fwd")
        gmean = torch.ones((), device="cuda")
        gpred = Guarded(n * hs * ws, tdt)
        L.check(lib.rsm_seqloss_bwd(gmean.data_ptr(), result.ptr(), pred.data_ptr(), gt.data_ptr(), valid.data_ptr(), gpred.ptr(),
                                    n, hs, ws, h, w, 700.0, kind, code, 0, st), "seqloss_bwd")
        gpred.check("rsm_seqloss_bwd")
    if (hs, ws) == (h, w):
        result = Guarded(8, torch.float64)
        L.check(lib.rsm_flow_metrics(gt.data_ptr(), pred.data_ptr(), valid.data_ptr(), work.data_ptr(), result.ptr(), n, 1, h, w,
                                     code, 0, st), "metrics")
        result.check("rsm_flow_metrics")
