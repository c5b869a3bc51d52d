"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's cost-volume and
disparity-regression algorithms (babiking/realtime_stereo_matcher, mounted read-only at
/root/reference in the build container).  Not product code; see ``oracle/__init__.py``.

Every function states the reference ``file:line`` it follows.  The arithmetic of the
reference lives in PyTorch ATen (torch 2.11.0+cu128 is the environment pin; the reference
has no requirements file), so float reductions are only order-equivalent, not
bit-equivalent, to ATen: copies, the difference volume, argmin and every reduction on
dyadic inputs (k/8) are bit-exact, the rest is pinned to the committed golden vectors
(``tests/golden``) within the tolerances written in ``tests/tolerances.py``.

Conventions: out-of-range entries (x < d) stay at the fill value and
take part in the regression; "soft-argmin" is softmax(+cost) expectation; layouts are
exactly the reference's and are never normalised.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "warp_by_flow_map", "warp_by_flow_map_bwd",
    "prepare_input", "prepare_input_bwd", "finalize_disparity", "finalize_disparity_bwd",
    "sequence_loss", "sequence_loss_bwd", "flow_map_metrics",
    "concat_volume", "concat_volume_bwd",
    "interweave", "interweave_bwd",
    "inner_product_volume", "inner_product_volume_bwd",
    "groupwise_volume", "groupwise_volume_bwd", "groupwise_pointwise",
    "difference_volume", "difference_volume_bwd", "shift_interweave_volume", "shift_interweave_volume_bwd",
    "softmax_d", "soft_argmax", "soft_argmax_bwd", "hard_argmin", "hard_argmax",
    "linear_axis_table", "trilinear_upsample", "trilinear_upsample_bwd",
    "v4_tail", "v4_tail_bwd", "inner_product_soft_argmax",
]


def _acc(a, acc_dtype):
    return a.astype(acc_dtype if acc_dtype is not None else a.dtype, copy=False)


# --------------------------------------------------------------------------- concatenate
def concat_volume(left, right, max_disparity):
    """(N,C,H,W)x2 -> (N,2C,H,W,D).  cost_volume/concatenate.py:11-41.

    V[:, :C, y, x, d] = L[:, :, y, x] and V[:, C:, y, x, d] = R[:, :, y, x-d] for x >= d,
    zero for x < d in both halves (concatenate.py:27-39).  d >= W leaves zeros.
    """
    n, c, h, w = left.shape
    vol = np.zeros((n, 2 * c, h, w, max_disparity), dtype=left.dtype)
    for d in range(min(max_disparity, w)):
        vol[:, :c, :, d:, d] = left[:, :, :, d:]
        vol[:, c:, :, d:, d] = right[:, :, :, : w - d]
    return vol


def concat_volume_bwd(gvol, acc_dtype=None):
    """Adjoint of :func:`concat_volume` (autograd through concatenate.py:33-39).

    gL[c,x] = sum_{d<=min(x,D-1)} gV[c,x,d];  gR[c,x'] = sum_{d<D, x'+d<W} gV[C+c,x'+d,d].
    """
    n, c2, h, w, dmax = gvol.shape
    c = c2 // 2
    g = _acc(gvol, acc_dtype)
    gl = np.zeros((n, c, h, w), dtype=g.dtype)
    gr = np.zeros((n, c, h, w), dtype=g.dtype)
    for d in range(min(dmax, w)):
        gl[:, :, :, d:] += g[:, :c, :, d:, d]
        gr[:, :, :, : w - d] += g[:, c:, :, d:, d]
    return gl.astype(gvol.dtype), gr.astype(gvol.dtype)


# ---------------------------------------------------------------------------- interweave
def interweave(left, right):
    """(N,C,H,W)x2 -> (N,2C,H,W), even channels = left, odd = right.

    cost_volume/interweave.py:10-22 and its twin interweave_tensors,
    model/mobile_stereo_net_v4.py:17-23.
    """
    n, c, h, w = left.shape
    out = np.zeros((n, 2 * c, h, w), dtype=left.dtype)
    out[:, 0::2] = left
    out[:, 1::2] = right
    return out


def interweave_bwd(gout):
    """gL = gV[:, 0::2], gR = gV[:, 1::2] (autograd through interweave.py:18-19)."""
    return np.ascontiguousarray(gout[:, 0::2]), np.ascontiguousarray(gout[:, 1::2])


# ------------------------------------------------------------- inner product / correlation
def inner_product_volume(left, right, max_disparity, mean=False, acc_dtype=np.float32,
                         out_dtype=None):
    """(N,C,H,W)x2 -> (N,D,H,W):  V[n,d,y,x] = s * sum_c L[n,c,y,x] R[n,c,y,x-d], x >= d.

    s = 1: TorchInnerProductCost.forward, cost_volume/inner_product.py:29-41 (channel sum).
    s = 1/C (``mean=True``): make_correlation_volume, model/mobile_disp_net_c.py:188-205
    (channel mean, applied as a division of the accumulated sum, like ``Tensor.mean``).
    Accumulates in ``acc_dtype`` (fp32 like ATen's reduction), output cast to
    ``out_dtype`` (default: dtype of ``left``).
    """
    n, c, h, w = left.shape
    out_dtype = out_dtype or left.dtype
    l = _acc(left, acc_dtype)
    r = _acc(right, acc_dtype)
    vol = np.zeros((n, max_disparity, h, w), dtype=l.dtype)
    for d in range(min(max_disparity, w)):
        s = np.sum(l[:, :, :, d:] * r[:, :, :, : w - d], axis=1)
        vol[:, d, :, d:] = s / c if mean else s
    return vol.astype(out_dtype)


def inner_product_volume_bwd(gvol, left, right, mean=False, acc_dtype=np.float32):
    """Adjoint of :func:`inner_product_volume` (SURVEY.md 8a "Backward contracts").

    gL[c,x] = s * sum_{d<=x} gV[d,x] R[c,x-d];  gR[c,x'] = s * sum_{d,x'+d<W} gV[d,x'+d] L[c,x'+d].
    """
    n, c, h, w = left.shape
    dmax = gvol.shape[1]
    g = _acc(gvol, acc_dtype)
    l = _acc(left, acc_dtype)
    r = _acc(right, acc_dtype)
    gl = np.zeros_like(l)
    gr = np.zeros_like(r)
    for d in range(min(dmax, w)):
        gd = g[:, d, :, d:][:, None]  # (N,1,H,W-d)
        gl[:, :, :, d:] += gd * r[:, :, :, : w - d]
        gr[:, :, :, : w - d] += gd * l[:, :, :, d:]
    if mean:
        gl = gl / c
        gr = gr / c
    return gl.astype(left.dtype), gr.astype(right.dtype)


# ----------------------------------------------------------------------------- groupwise
def groupwise_pointwise(left, right, n_groups, acc_dtype=np.float32):
    """(N,C,H,W)x2 -> (N,G,H,W): per-group channel mean of L*R.

    TorchGroupwiseCost.groupwise, cost_volume/groupwise.py:12-22; groups are contiguous
    channel blocks of C//G; AssertionError when C % G != 0 (groupwise.py:15-17).
    """
    n, c, h, w = left.shape
    assert c % n_groups == 0, f"groupwise cost channel ({c}) % #groups ({n_groups}) != 0."
    cpg = c // n_groups
    prod = _acc(left, acc_dtype) * _acc(right, acc_dtype)
    return prod.reshape(n, n_groups, cpg, h, w).sum(axis=2) / cpg


def groupwise_volume(left, right, n_groups, max_disparity, acc_dtype=np.float32,
                     out_dtype=np.float32):
    """(N,C,H,W)x2 -> (N,G,H,W,D).  cost_volume/groupwise.py:24-56.

    V[n,g,y,x,d] = (1/cpg) sum_{c in g} L[n,c,y,x] R[n,c,y,x-d], x >= d, zero elsewhere.
    The reference allocates the output without dtype/device (groupwise.py:39), so it is
    always fp32; ``out_dtype`` defaults to that.
    """
    n, c, h, w = left.shape
    vol = np.zeros((n, n_groups, h, w, max_disparity), dtype=out_dtype)
    for d in range(min(max_disparity, w)):
        vol[:, :, :, d:, d] = groupwise_pointwise(
            left[:, :, :, d:], right[:, :, :, : w - d], n_groups, acc_dtype)
    return vol


def groupwise_volume_bwd(gvol, left, right, n_groups, acc_dtype=np.float32):
    """Adjoint of :func:`groupwise_volume`: inner-product adjoint with per-group gV, s=1/cpg."""
    n, c, h, w = left.shape
    dmax = gvol.shape[-1]
    cpg = c // n_groups
    g = _acc(gvol, acc_dtype)
    l = _acc(left, acc_dtype)
    r = _acc(right, acc_dtype)
    gl = np.zeros_like(l)
    gr = np.zeros_like(r)
    for d in range(min(dmax, w)):
        gd = np.repeat(g[:, :, :, d:, d], cpg, axis=1)  # (N,C,H,W-d)
        gl[:, :, :, d:] += gd * r[:, :, :, : w - d]
        gr[:, :, :, : w - d] += gd * l[:, :, :, d:]
    return (gl / cpg).astype(left.dtype), (gr / cpg).astype(right.dtype)


# ---------------------------------------------------------------------------- difference
def difference_volume(left, right, max_disp):
    """(N,C,H,W)x2 -> (N,C,D,H,W): L - R shifted by d, fill value 1.0 for x < d.

    make_cost_volume, model/mobile_stereo_net.py:8-27 (identical copies in
    model/mobile_stereo_net_v2.py:8-27 and model/mobile_stereo_net_v3.py:9-28).
    """
    n, c, h, w = left.shape
    vol = np.ones((n, c, max_disp, h, w), dtype=left.dtype)
    for d in range(min(max_disp, w)):
        vol[:, :, d, :, d:] = left[:, :, :, d:] - right[:, :, :, : w - d]
    return vol


def difference_volume_bwd(gvol, acc_dtype=None):
    """gL[c,x] = sum_{d<=x} gV[c,d,x];  gR[c,x'] = -sum_{d,x'+d<W} gV[c,d,x'+d]."""
    n, c, dmax, h, w = gvol.shape
    g = _acc(gvol, acc_dtype)
    gl = np.zeros((n, c, h, w), dtype=g.dtype)
    gr = np.zeros((n, c, h, w), dtype=g.dtype)
    for d in range(min(dmax, w)):
        gl[:, :, :, d:] += g[:, :, d, :, d:]
        gr[:, :, :, : w - d] -= g[:, :, d, :, d:]
    return gl.astype(gvol.dtype), gr.astype(gvol.dtype)


# ------------------------------------------------------------- shifted interweave stack
def shift_interweave_volume(left, right, max_disparity):
    """(N,C,H,W)x2 -> (D,N,2C,H,W): out[d] = interweave(left, right shifted right by d), zero for x < d.

    Iteration i of the MobileStereoNetV4 volume loop (model/mobile_stereo_net_v4.py:444-458) feeds
    interweave_tensors(featL[..., i:], featR[..., :-i]) to its convolutions; out[i][..., i:] is exactly that
    tensor and out[i][..., :i] = 0 reproduces the zero padding those cropped convolutions see."""
    n, c, h, w = left.shape
    out = np.zeros((max_disparity, n, 2 * c, h, w), dtype=left.dtype)
    for d in range(min(max_disparity, w)):
        out[d, :, 0::2, :, d:] = left[:, :, :, d:]
        out[d, :, 1::2, :, d:] = right[:, :, :, : w - d]
    return out


def shift_interweave_volume_bwd(gout, acc_dtype=None):
    """gL[c,x] = sum_{d<=x} g[d,2c,x];  gR[c,x'] = sum_{d, x'+d<W} g[d,2c+1,x'+d]."""
    dmax, n, c2, h, w = gout.shape
    g = _acc(gout, acc_dtype)
    gl = np.zeros((n, c2 // 2, h, w), dtype=g.dtype)
    gr = np.zeros_like(gl)
    for d in range(min(dmax, w)):
        gl[:, :, :, d:] += g[d, :, 0::2, :, d:]
        gr[:, :, :, : w - d] += g[d, :, 1::2, :, d:]
    return gl.astype(gout.dtype), gr.astype(gout.dtype)


# ---------------------------------------------------------------------------- regression
def softmax_d(cost, acc_dtype=np.float32):
    """softmax over axis 1 (the disparity axis), F.softmax(cost, dim=1) as used at
    model/mobile_stereo_net.py:144, mobile_stereo_net_v4.py:517, mobile_disp_net_c.py:218."""
    c = _acc(cost, acc_dtype)
    m = np.max(c, axis=1, keepdims=True)
    e = np.exp(c - m)
    return e / np.sum(e, axis=1, keepdims=True)


def soft_argmax(cost, acc_dtype=np.float32, keepdim=False):
    """(N,D,H,W) -> (N,H,W) [or (N,1,H,W)]:  E = sum_d d * softmax_d(+cost).

    Inline form model/mobile_stereo_net.py:144-147 (= v2 :217-220, v3 :321-324);
    disparity_regression of probabilities model/mobile_stereo_net_v4.py:10-14;
    disparity_regression of logits model/mobile_disp_net_c.py:208-220.
    Softmax of +cost, fill-value entries take part (F8).
    """
    p = softmax_d(cost, acc_dtype)
    d = np.arange(cost.shape[1], dtype=p.dtype).reshape(1, -1, 1, 1)
    e = np.sum(p * d, axis=1, keepdims=keepdim)
    return e.astype(cost.dtype)


def soft_argmax_bwd(gout, cost, acc_dtype=np.float32):
    """gcost[d] = g * p[d] * (d - E) (softmax + expectation adjoint)."""
    p = softmax_d(cost, acc_dtype)
    d = np.arange(cost.shape[1], dtype=p.dtype).reshape(1, -1, 1, 1)
    e = np.sum(p * d, axis=1, keepdims=True)
    g = _acc(gout, acc_dtype).reshape(cost.shape[0], 1, cost.shape[2], cost.shape[3])
    return (g * p * (d - e)).astype(cost.dtype)


def hard_argmin(cost):
    """(N,D,H,W) -> (N,H,W) int64 = torch.argmin(cost, dim=1) semantics.

    Not in the reference: first index on ties, NaN counts as the minimum
    (first NaN wins), int64 result.  numpy.argmin has the same rules.
    """
    return np.argmin(cost, axis=1).astype(np.int64)


def hard_argmax(cost):
    """torch.argmax(cost, dim=1): first index on ties, first NaN wins, int64."""
    return np.argmax(cost, axis=1).astype(np.int64)


# ------------------------------------------------------------------------- v4 tail
def linear_axis_table(n_in, n_out):
    """Source indices and weights of F.interpolate(mode='trilinear', align_corners=False)
    along one axis, as ATen computes them in fp32: scale = in/out,
    src = max(scale*(dst+0.5)-0.5, 0), i0 = floor(src), i1 = i0 + (i0 < in-1), w1 = src-i0.
    Used by model/mobile_stereo_net_v4.py:513-516 (and :476-506 in training).
    """
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = scale * (dst + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    w1 = np.clip(src - i0.astype(np.float32), 0.0, 1.0).astype(np.float32)
    w0 = (np.float32(1.0) - w1).astype(np.float32)
    return i0, i1, w0, w1


def trilinear_upsample(cost, out_d, out_h, out_w, acc_dtype=np.float32):
    """(N,Dc,Hc,Wc) -> (N,out_d,out_h,out_w): F.interpolate(cost[:,None], [D,H,W],
    mode='trilinear') squeezed, model/mobile_stereo_net_v4.py:512-516 (separable lerps)."""
    c = _acc(cost, acc_dtype)
    n, dc, hc, wc = c.shape
    x0, x1, wx0, wx1 = linear_axis_table(wc, out_w)
    y0, y1, wy0, wy1 = linear_axis_table(hc, out_h)
    d0, d1, wd0, wd1 = linear_axis_table(dc, out_d)
    wx0, wx1, wy0, wy1, wd0, wd1 = (a.astype(c.dtype) for a in (wx0, wx1, wy0, wy1, wd0, wd1))
    t = c[:, :, :, x0] * wx0 + c[:, :, :, x1] * wx1                       # (N,Dc,Hc,W)
    t = t[:, :, y0, :] * wy0[:, None] + t[:, :, y1, :] * wy1[:, None]     # (N,Dc,H,W)
    t = t[:, d0] * wd0[:, None, None] + t[:, d1] * wd1[:, None, None]     # (N,D,H,W)
    return t


def trilinear_upsample_bwd(gfine, dc, hc, wc, acc_dtype=np.float32):
    """Transpose of :func:`trilinear_upsample` (UpsampleTrilinear3DBackward)."""
    g = _acc(gfine, acc_dtype)
    n, od, oh, ow = g.shape
    x0, x1, wx0, wx1 = linear_axis_table(wc, ow)
    y0, y1, wy0, wy1 = linear_axis_table(hc, oh)
    d0, d1, wd0, wd1 = linear_axis_table(dc, od)
    wx0, wx1, wy0, wy1, wd0, wd1 = (a.astype(g.dtype) for a in (wx0, wx1, wy0, wy1, wd0, wd1))
    td = np.zeros((n, dc, oh, ow), dtype=g.dtype)
    np.add.at(td, (slice(None), d0), g * wd0[:, None, None])
    np.add.at(td, (slice(None), d1), g * wd1[:, None, None])
    ty = np.zeros((n, dc, hc, ow), dtype=g.dtype)
    np.add.at(ty, (slice(None), slice(None), y0), td * wy0[:, None])
    np.add.at(ty, (slice(None), slice(None), y1), td * wy1[:, None])
    tx = np.zeros((n, dc, hc, wc), dtype=g.dtype)
    np.add.at(tx, (slice(None), slice(None), slice(None), x0), ty * wx0)
    np.add.at(tx, (slice(None), slice(None), slice(None), x1), ty * wx1)
    return tx


def v4_tail(cost, maxdisp, out_h, out_w, acc_dtype=np.float32):
    """(B,Dc,Hc,Wc) -> (B,out_h,out_w): trilinear upsample to (maxdisp,H,W), softmax over
    D, expectation.  Eval head model/mobile_stereo_net_v4.py:511-518 (training heads
    :471-506 are four copies).  The caller negates (:520)."""
    fine = trilinear_upsample(cost, maxdisp, out_h, out_w, acc_dtype)
    return soft_argmax(fine.astype(np.float32 if acc_dtype == np.float32 else acc_dtype),
                       acc_dtype).astype(cost.dtype)


def v4_tail_bwd(gout, cost, maxdisp, out_h, out_w, acc_dtype=np.float32):
    """Adjoint of :func:`v4_tail` with respect to the coarse cost."""
    fine = trilinear_upsample(cost, maxdisp, out_h, out_w, acc_dtype)
    gfine = soft_argmax_bwd(gout, fine, acc_dtype)
    n, dc, hc, wc = cost.shape
    return trilinear_upsample_bwd(gfine, dc, hc, wc, acc_dtype).astype(cost.dtype)


def inner_product_soft_argmax(left, right, max_disparity, mean=False, acc_dtype=np.float32):
    """Composition used by the fused no-volume kernel: inner_product_volume followed by
    soft_argmax and hard_argmax/argmin on the fp32 volume (never rounded to the input dtype)."""
    vol = inner_product_volume(left, right, max_disparity, mean, acc_dtype, out_dtype=acc_dtype)
    return soft_argmax(vol, acc_dtype), hard_argmin(vol), hard_argmax(vol)


# ------------------------------------------------------------------ refinement warp
def _warp_source_index(flow, h, w):
    """fp32 source coordinates exactly as the reference builds them: grid = index - flow, normalised with
    (size - 1) (model/mobile_stereo_net_v2.py:82-91) but sampled by F.grid_sample(align_corners=False)
    (:93-95), whose un-normalisation is ((g + 1) * size - 1) / 2 -- so ix = (x - f) * W / (W - 1) - 0.5 and,
    even for a 1-channel flow, iy = y * H / (H - 1) - 0.5 is NOT an integer row (the reference's quirk, kept)."""
    f = np.asarray(flow, dtype=np.float32)
    n = f.shape[0]
    one, two = np.float32(1), np.float32(2)
    gx = np.arange(w, dtype=np.float32)[None, None, :] - f[:, 0]
    gy = (np.arange(h, dtype=np.float32)[None, :, None] - f[:, 1]) if f.shape[1] == 2 else \
        np.broadcast_to(np.arange(h, dtype=np.float32)[None, :, None], (n, h, w)).astype(np.float32)
    gxn = two * gx / np.float32(w - 1.0) - one
    gyn = two * gy / np.float32(h - 1.0) - one
    ix = ((gxn + one) * np.float32(w) - one) / two
    iy = ((gyn + one) * np.float32(h) - one) / two
    return ix.astype(np.float32), iy.astype(np.float32)


def _warp_taps(ix, iy, h, w):
    """the four bilinear taps of grid_sample(padding_mode='zeros'): (yi, xi, weight, inside) per corner"""
    x0 = np.floor(ix); y0 = np.floor(iy)
    taps = []
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx; yi = y0 + dy
            wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
            wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
            inside = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            taps.append((np.clip(yi, 0, h - 1).astype(np.int64), np.clip(xi, 0, w - 1).astype(np.int64),
                         (wx * wy).astype(np.float32), inside, dx, dy, wx.astype(np.float32), wy.astype(np.float32)))
    return taps


def warp_by_flow_map(image, flow):
    """(N,C,H,W), (N,1|2,H,W) -> (N,C,H,W): warp_by_flow_map, model/mobile_stereo_net_v2.py:59-96
    (= model/mobile_stereo_net_v3.py:60-97, tools/warp.py:5-42)."""
    img = np.asarray(image)
    n, c, h, w = img.shape
    cf = flow.shape[1]
    assert cf == 1 or cf == 2, f"invalid flow map dimension 1 or 2 ({cf})!"
    ix, iy = _warp_source_index(flow, h, w)
    out = np.zeros((n, c, h, w), dtype=np.float32)
    bi = np.arange(n)[:, None, None]
    for yi, xi, wt, inside, *_ in _warp_taps(ix, iy, h, w):
        v = img[bi, :, yi, xi].astype(np.float32)              # (N,H,W,C)
        out += np.moveaxis(v * (wt * inside)[..., None], -1, 1)
    return out.astype(img.dtype)


def warp_by_flow_map_bwd(gout, image, flow):
    """adjoint of :func:`warp_by_flow_map`: (gimage, gflow).  d ix / d f0 = -W / (W - 1), d iy / d f1 = -H / (H - 1)."""
    img = np.asarray(image, dtype=np.float32)
    g = np.asarray(gout, dtype=np.float32)
    n, c, h, w = img.shape
    cf = flow.shape[1]
    ix, iy = _warp_source_index(flow, h, w)
    gimg = np.zeros((n, c, h, w), dtype=np.float64)
    gix = np.zeros((n, h, w), dtype=np.float64)
    giy = np.zeros((n, h, w), dtype=np.float64)
    bi = np.broadcast_to(np.arange(n)[:, None, None], (n, h, w))
    for yi, xi, wt, inside, dx, dy, wx, wy in _warp_taps(ix, iy, h, w):
        wgt = (wt * inside).astype(np.float64)
        for ch in range(c):
            np.add.at(gimg[:, ch], (bi, yi, xi), g[:, ch] * wgt)
        v = np.moveaxis(img[bi, :, yi, xi], -1, 1) * inside[:, None]          # (N,C,H,W), zero outside
        sx = 1.0 if dx == 1 else -1.0
        sy = 1.0 if dy == 1 else -1.0
        gix += (g * v).sum(1) * sx * wy
        giy += (g * v).sum(1) * sy * wx
    gflow = np.zeros(flow.shape, dtype=np.float64)
    gflow[:, 0] = -gix * (w / (w - 1.0))
    if cf == 2:
        gflow[:, 1] = -giy * (h / (h - 1.0))
    return gimg.astype(np.asarray(image).dtype), gflow.astype(np.asarray(flow).dtype)


# ------------------------------------------------------------------ pre / post steps
def _div255(x, device_div):
    """``x / 255.0`` as torch evaluates it: a true division on the CPU; on a CUDA tensor ATen's div-by-scalar kernel
    multiplies by the reciprocal rounded to fp32 once (BinaryDivTrueKernel.cu), computing in fp32 for 16-bit tensors."""
    dt = x.dtype.type
    if not device_div:
        return x / dt(255.0)
    return (x.astype(np.float32) * (np.float32(1.0) / np.float32(255.0))).astype(x.dtype)


def prepare_input(img, align=1, device_div=False):
    """(N,C,H,W) -> (N,C,Hp,Wp): model/mobile_stereo_net.py:121-130 (= _v2.py:194-203, _v3.py:296-305,
    mobile_disp_net_c.py:339-351): 2 * (img / 255) - 1 in the tensor's dtype, then zero padding on the right /
    bottom to a multiple of ``align``.  ``device_div``: the division as torch's CUDA kernel performs it (what the
    reference computes on a GPU, and what rsm_prepare_fwd follows); default = the CPU's true division (goldens)."""
    img = np.asarray(img)
    dt = img.dtype.type
    x = dt(2.0) * _div255(img, device_div) - dt(1.0)
    h, w = img.shape[2:]
    h_pad = (align - (h % align)) % align
    w_pad = (align - (w % align)) % align
    return np.pad(x, ((0, 0), (0, 0), (0, h_pad), (0, w_pad)))


def prepare_input_bwd(gout, size, device_div=False):
    """adjoint of prepare_input: (gout[:, :, :H, :W] * 2) / 255."""
    gout = np.asarray(gout)
    dt = gout.dtype.type
    h, w = size
    return _div255(gout[:, :, :h, :w] * dt(2.0), device_div)


def _resize_taps(out_size, in_size, mode):
    """ATen's source indices / weights for one axis (UpSample.h): scale = float(in) / out;
    nearest: min(int(floorf(dst * scale)), in - 1); linear (align_corners=False): src = max(scale * (dst + 0.5) - 0.5, 0),
    i0 = int(src), i1 = i0 + (i0 < in - 1), l1 = src - i0."""
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    if mode == "nearest":
        i0 = np.minimum(np.floor(dst * scale).astype(np.int64), in_size - 1)
        return i0, i0, np.ones(out_size, f32), np.zeros(out_size, f32)
    src = np.maximum(scale * (dst + f32(0.5)) - f32(0.5), f32(0.0)).astype(f32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.clip(src - i0.astype(f32), 0.0, 1.0).astype(f32)
    return i0, i1, (f32(1.0) - l1).astype(f32), l1


def _final_scale(disp_shape, padded_size, mode):
    hs, ws = disp_shape[2:]
    if mode == "bilinear" and (hs, ws) == tuple(padded_size):
        return np.float32(1.0)          # disparity_interpolate: same-size maps pass through (mobile_disp_net_c.py:228)
    return np.float32(float(padded_size[1]) / ws)


def finalize_disparity(disp, padded_size, size=None, mode="nearest"):
    """(N,C,hs,ws) -> (N,C,h,w): -1.0 * F.interpolate(disp * scale, padded_size)[:, :, :h, :w], scale = Wp / ws --
    model/mobile_stereo_net.py:154-159 (= _v2.py:227-232, nearest) and model/mobile_disp_net_c.py:223-234 + :408-411
    (bilinear, align_corners=False)."""
    disp = np.asarray(disp)
    hp, wp = padded_size
    h, w = (hp, wp) if size is None else size
    hs, ws = disp.shape[2:]
    v = (disp.astype(np.float32) * _final_scale(disp.shape, padded_size, mode)).astype(disp.dtype).astype(np.float32)
    y0, y1, ly0, ly1 = (a[:h] for a in _resize_taps(hp, hs, mode))
    x0, x1, lx0, lx1 = (a[:w] for a in _resize_taps(wp, ws, mode))
    if mode == "nearest":
        out = v[:, :, y0][:, :, :, x0]
    else:
        top = lx0 * v[:, :, y0][:, :, :, x0] + lx1 * v[:, :, y0][:, :, :, x1]
        bot = lx0 * v[:, :, y1][:, :, :, x0] + lx1 * v[:, :, y1][:, :, :, x1]
        out = ly0[:, None] * top + ly1[:, None] * bot
    return (-out).astype(disp.dtype)


def finalize_disparity_bwd(gout, disp_shape, padded_size, mode="nearest"):
    """adjoint of finalize_disparity with respect to disp (float64 accumulation)."""
    gout = np.asarray(gout)
    n, c, h, w = gout.shape
    hs, ws = disp_shape[2:]
    hp, wp = padded_size
    y0, y1, ly0, ly1 = (a[:h] for a in _resize_taps(hp, hs, mode))
    x0, x1, lx0, lx1 = (a[:w] for a in _resize_taps(wp, ws, mode))
    my = np.zeros((h, hs)); mx = np.zeros((w, ws))
    np.add.at(my, (np.arange(h), y0), ly0); np.add.at(my, (np.arange(h), y1), ly1)
    np.add.at(mx, (np.arange(w), x0), lx0); np.add.at(mx, (np.arange(w), x1), lx1)
    g = np.einsum("ys,ncyx,xt->ncst", my, gout.astype(np.float64), mx)
    return (-float(_final_scale(disp_shape, padded_size, mode)) * g).astype(gout.dtype)


# ------------------------------------------------------------------ loss / metrics
def _loss_mask(flow_gt, flow_valid, max_flow):
    """loss/loss.py:55-58: (flow_valid >= 0.5) & (sqrt(sum_c gt^2) < max_flow), shape (N,1,H,W)."""
    mag = np.sqrt(np.sum(flow_gt.astype(np.float32) ** 2, axis=1, dtype=np.float32))
    return ((flow_valid >= 0.5) & (mag < max_flow))[:, None]


def _loss_resize(pred, size):
    """loss/loss.py:71-73: F.interpolate(pred * (W / w), size) (nearest) when the shapes differ."""
    hs, ws = pred.shape[2:]
    if (hs, ws) == tuple(size):
        return pred.astype(np.float32), np.float32(1.0), None, None
    scale = np.float32(float(size[1]) / ws)
    y0 = _resize_taps(size[0], hs, "nearest")[0]
    x0 = _resize_taps(size[1], ws, "nearest")[0]
    return (pred.astype(np.float32) * scale)[:, :, y0][:, :, :, x0], scale, y0, x0


def sequence_loss(flow_preds, flow_gt, flow_valid, loss_gamma=0.9, max_flow=700):
    """SequenceLoss.forward, loss/loss.py:36-81: sum_i gamma^(n-1-i) * mean over the valid pixels of |gt - p_i|
    (smooth-L1 with beta 1 for the last prediction); float64 accumulation."""
    mask = _loss_mask(flow_gt, flow_valid, max_flow)
    n = len(flow_preds)
    total = 0.0
    for i, pred in enumerate(flow_preds):
        p = _loss_resize(np.asarray(pred), flow_gt.shape[2:])[0]
        d = np.abs(flow_gt.astype(np.float32) - p)
        el = np.where(d < 1.0, np.float32(0.5) * d * d, d - np.float32(0.5)) if i == n - 1 else d
        total += loss_gamma ** (n - 1 - i) * el[mask].astype(np.float64).mean()
    return np.float32(total)


def sequence_loss_bwd(flow_preds, flow_gt, flow_valid, loss_gamma=0.9, max_flow=700, gout=1.0):
    """gradients of sequence_loss with respect to every prediction."""
    mask = _loss_mask(flow_gt, flow_valid, max_flow)
    count = float(mask.sum())
    n = len(flow_preds)
    grads = []
    for i, pred in enumerate(flow_preds):
        pred = np.asarray(pred)
        p, scale, y0, x0 = _loss_resize(pred, flow_gt.shape[2:])
        d = flow_gt.astype(np.float32) - p
        if i == n - 1:
            g = np.where(np.abs(d) < 1.0, -d, -np.sign(d))
        else:
            g = -np.sign(d)
        g = np.where(mask, g, 0.0).astype(np.float64) * (gout * loss_gamma ** (n - 1 - i) / count)
        if y0 is None:
            grads.append(g.astype(pred.dtype))
            continue
        hs, ws = pred.shape[2:]
        acc = np.zeros(pred.shape, np.float64)
        np.add.at(acc, (slice(None), slice(None), y0[:, None], x0[None, :]), g)
        grads.append((acc * float(scale)).astype(pred.dtype))
    return grads


def flow_map_metrics(flow_gt, flow_pred, flow_valid):
    """get_flow_map_metrics, loss/loss.py:6-22."""
    epe = np.sqrt(np.sum((flow_pred.astype(np.float32) - flow_gt.astype(np.float32)) ** 2, axis=1, dtype=np.float32))
    epe = epe.reshape(-1)[(flow_valid >= 0.5).reshape(-1)]
    return {"epe": float(epe.astype(np.float64).mean()), "0.5px": float((epe < 0.5).mean()), "1px": float((epe < 1).mean()),
            "3px": float((epe < 3).mean()), "5px": float((epe < 5).mean()),
            "min": float(flow_pred[0].min()), "max": float(flow_pred[0].max())}
