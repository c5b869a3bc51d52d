"""Functional form of the hot path: every op is one C-ABI call into librsm_b200.so, with forward
and backward wired into autograd so the reference's train_stereo.py keeps working.

Shapes/layouts are the reference's: concat (N,2C,H,W,D), interweave (N,2C,H,W),
inner/correlation (N,D,H,W), groupwise (N,G,H,W,D), difference (N,C,D,H,W).  Out-of-range
entries (x < d) hold the fill value and take part in the regression (F8); the regression is
softmax(+cost) (F9).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


def _check_pair(left: torch.Tensor, right: torch.Tensor) -> Tuple[int, int, int, int, int]:
    if left.dim() != 4 or right.dim() != 4:
        raise ValueError(f"expected (N,C,H,W) feature maps, got {tuple(left.shape)} and {tuple(right.shape)}")
    if left.shape != right.shape:
        raise RuntimeError(f"left {tuple(left.shape)} and right {tuple(right.shape)} feature shapes differ")
    if left.dtype != right.dtype:
        raise TypeError(f"left ({left.dtype}) and right ({right.dtype}) dtypes differ")
    dev = L.require_cuda(left, right)
    n, c, h, w = left.shape
    return dev, n, c, h, w


def _dense(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------- concatenate
class _Concat(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right, max_disparity):
        dev, n, c, h, w = _check_pair(left, right)
        d = int(max_disparity)
        out = torch.empty((n, 2 * c, h, w, d), dtype=left.dtype, device=left.device)
        L.check(L.load().rsm_concat_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w, d,
                                        L.dtype_code(left), dev, L.stream_ptr(dev)), "rsm_concat_fwd")
        ctx.dims = (dev, n, c, h, w, d)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, h, w, d = ctx.dims
        gout = _dense(gout)
        gl = torch.empty((n, c, h, w), dtype=gout.dtype, device=gout.device)
        gr = torch.empty_like(gl)
        L.check(L.load().rsm_concat_bwd(gout.data_ptr(), gl.data_ptr(), gr.data_ptr(), n, c, h, w, d,
                                        L.dtype_code(gout), dev, L.stream_ptr(dev)), "rsm_concat_bwd")
        return gl, gr, None


def concat_volume(left, right, max_disparity):
    """TorchConcatenateCost.forward (reference cost_volume/concatenate.py:11-41)."""
    return _Concat.apply(left, right, max_disparity)


# ---------------------------------------------------------------------------- interweave
class _Interweave(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right):
        dev, n, c, h, w = _check_pair(left, right)
        out = torch.empty((n, 2 * c, h, w), dtype=left.dtype, device=left.device)
        L.check(L.load().rsm_interweave_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w,
                                            L.dtype_code(left), dev, L.stream_ptr(dev)), "rsm_interweave_fwd")
        ctx.dims = (dev, n, c, h, w)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, h, w = ctx.dims
        gout = _dense(gout)
        gl = torch.empty((n, c, h, w), dtype=gout.dtype, device=gout.device)
        gr = torch.empty_like(gl)
        L.check(L.load().rsm_interweave_bwd(gout.data_ptr(), gl.data_ptr(), gr.data_ptr(), n, c, h, w,
                                            L.dtype_code(gout), dev, L.stream_ptr(dev)), "rsm_interweave_bwd")
        return gl, gr


def interweave(left, right):
    """TorchInterweaveCost.forward / interweave_tensors (interweave.py:10-22, mobile_stereo_net_v4.py:17-23)."""
    return _Interweave.apply(left, right)


# ------------------------------------------------------------- inner product / correlation
def _out_dtype(left: torch.Tensor, out_dtype: Optional[torch.dtype]) -> torch.dtype:
    out_dtype = out_dtype or left.dtype
    if out_dtype not in (left.dtype, torch.float32):
        raise TypeError(f"out_dtype must be the input dtype or float32, got {out_dtype}")
    return out_dtype


class _Inner(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right, max_disparity, mean, out_dtype):
        dev, n, c, h, w = _check_pair(left, right)
        d = int(max_disparity)
        odt = _out_dtype(left, out_dtype)
        out = torch.empty((n, d, h, w), dtype=odt, device=left.device)
        red = L.RSM_REDUCE_MEAN if mean else L.RSM_REDUCE_SUM
        L.check(L.load().rsm_inner_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w, d, red,
                                       L.dtype_code(left), L.dtype_code(out), dev, L.stream_ptr(dev)),
                "rsm_inner_fwd")
        ctx.save_for_backward(left, right)
        ctx.dims = (dev, n, c, h, w, d, red, odt)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        left, right = ctx.saved_tensors
        dev, n, c, h, w, d, red, odt = ctx.dims
        gout = _dense(gout.to(odt))
        gl = torch.empty((n, c, h, w), dtype=left.dtype, device=left.device) if ctx.needs_input_grad[0] else None
        gr = torch.empty((n, c, h, w), dtype=left.dtype, device=left.device) if ctx.needs_input_grad[1] else None
        L.check(L.load().rsm_inner_bwd(gout.data_ptr(), L.feat(left), L.feat(right), L.ptr(gl), L.ptr(gr),
                                       n, c, h, w, d, red, L.dtype_code(left), L.dtype_code(gout), dev,
                                       L.stream_ptr(dev)), "rsm_inner_bwd")
        return gl, gr, None, None, None


def inner_product_volume(left, right, max_disparity, mean=False, out_dtype=None):
    """TorchInnerProductCost.forward (inner_product.py:11-42; channel sum) or, with ``mean=True``,
    make_correlation_volume (mobile_disp_net_c.py:188-205; channel mean).  fp32 accumulation."""
    return _Inner.apply(left, right, max_disparity, bool(mean), out_dtype)


# ----------------------------------------------------------------------------- groupwise
class _Groupwise(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right, n_groups, max_disparity, out_dtype):
        dev, n, c, h, w = _check_pair(left, right)
        g, d = int(n_groups), int(max_disparity)
        # same check and message as the reference (cost_volume/groupwise.py:15-17)
        assert c % g == 0, f"groupwise cost channel ({c}) % #groups ({g}) != 0."
        odt = _out_dtype(left, out_dtype)
        out = torch.empty((n, g, h, w, d), dtype=odt, device=left.device)
        L.check(L.load().rsm_groupwise_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w, d, g,
                                           L.dtype_code(left), L.dtype_code(out), dev, L.stream_ptr(dev)),
                "rsm_groupwise_fwd")
        ctx.save_for_backward(left, right)
        ctx.dims = (dev, n, c, h, w, d, g, odt)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        left, right = ctx.saved_tensors
        dev, n, c, h, w, d, g, odt = ctx.dims
        gout = _dense(gout.to(odt))
        gl = torch.empty((n, c, h, w), dtype=left.dtype, device=left.device) if ctx.needs_input_grad[0] else None
        gr = torch.empty((n, c, h, w), dtype=left.dtype, device=left.device) if ctx.needs_input_grad[1] else None
        L.check(L.load().rsm_groupwise_bwd(gout.data_ptr(), L.feat(left), L.feat(right), L.ptr(gl), L.ptr(gr),
                                           n, c, h, w, d, g, L.dtype_code(left), L.dtype_code(gout), dev,
                                           L.stream_ptr(dev)), "rsm_groupwise_bwd")
        return gl, gr, None, None, None


def groupwise_volume(left, right, n_groups, max_disparity, out_dtype=None):
    """TorchGroupwiseCost.forward (groupwise.py:24-56).  The volume is returned on ``left.device``
    in ``left.dtype`` (pass ``out_dtype=torch.float32`` for the reference's always-fp32 output;
    the reference's CPU placement is its bug)."""
    return _Groupwise.apply(left, right, n_groups, max_disparity, out_dtype)


def groupwise_pointwise(left, right, n_groups):
    """TorchGroupwiseCost.groupwise (groupwise.py:12-22): (N,G,H,W) = the d=0 slice."""
    return groupwise_volume(left, right, n_groups, 1)[..., 0]


# ---------------------------------------------------------------------------- difference
class _Difference(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right, max_disp, fill):
        dev, n, c, h, w = _check_pair(left, right)
        d = int(max_disp)
        out = torch.empty((n, c, d, h, w), dtype=left.dtype, device=left.device)
        L.check(L.load().rsm_difference_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w, d,
                                            float(fill), L.dtype_code(left), dev, L.stream_ptr(dev)),
                "rsm_difference_fwd")
        ctx.dims = (dev, n, c, h, w, d)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, h, w, d = ctx.dims
        gout = _dense(gout)
        gl = torch.empty((n, c, h, w), dtype=gout.dtype, device=gout.device)
        gr = torch.empty_like(gl)
        L.check(L.load().rsm_difference_bwd(gout.data_ptr(), gl.data_ptr(), gr.data_ptr(), n, c, h, w, d,
                                            L.dtype_code(gout), dev, L.stream_ptr(dev)), "rsm_difference_bwd")
        return gl, gr, None, None


def difference_volume(left, right, max_disp, fill=1.0):
    """make_cost_volume of MobileStereoNet v1-v3 (mobile_stereo_net.py:8-27): L - R shifted, fill 1.0."""
    return _Difference.apply(left, right, max_disp, fill)


# ------------------------------------------------------------- shifted interweave stack
class _ShiftInterweave(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, left, right, max_disparity):
        dev, n, c, h, w = _check_pair(left, right)
        d = int(max_disparity)
        out = torch.empty((d, n, 2 * c, h, w), dtype=left.dtype, device=left.device)
        L.check(L.load().rsm_shift_interweave_fwd(L.feat(left), L.feat(right), out.data_ptr(), n, c, h, w, d,
                                                  L.dtype_code(left), dev, L.stream_ptr(dev)),
                "rsm_shift_interweave_fwd")
        ctx.dims = (dev, n, c, h, w, d)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, h, w, d = ctx.dims
        gout = _dense(gout)
        gl = torch.empty((n, c, h, w), dtype=gout.dtype, device=gout.device)
        gr = torch.empty_like(gl)
        L.check(L.load().rsm_shift_interweave_bwd(gout.data_ptr(), gl.data_ptr(), gr.data_ptr(), n, c, h, w, d,
                                                  L.dtype_code(gout), dev, L.stream_ptr(dev)),
                "rsm_shift_interweave_bwd")
        return gl, gr, None


def shift_interweave_volume(left, right, max_disparity):
    """All iterations of MobileStereoNetV4's per-disparity loop input at once (mobile_stereo_net_v4.py:444-458):
    out (D,N,2C,H,W), out[d] = interweave_tensors(left, right shifted by d), zero where x < d."""
    return _ShiftInterweave.apply(left, right, max_disparity)


# ------------------------------------------------------------- v4 learned per-disparity volume
_V4_WEIGHT_CACHE = {}


def _fold_bn(conv_w, conv_b, bn):
    """Eval-mode BatchNorm folded into the convolution before it: scale per output channel + additive term."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    b = conv_b.detach().float() if conv_b is not None else torch.zeros_like(s)
    t = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return conv_w.detach().float() * s.view(-1, *([1] * (conv_w.dim() - 1))), t


def pack_v4_weights(conv3d, volume11, op_dtype):
    """Fold + pack the weights of MobileStereoNetV4.conv3d / .volume11 (mobile_stereo_net_v4.py:317-335) for
    rsm_v4_volume_fwd (layouts in include/rsm.h).  Cached per module pair until a parameter / buffer changes."""
    mods = [conv3d[0], conv3d[1], conv3d[3], conv3d[4], conv3d[6], conv3d[7], volume11[0][0], volume11[0][1]]
    tensors = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
    key = (id(conv3d), id(volume11), op_dtype, tuple((t.data_ptr(), t._version) for t in tensors))
    hit = _V4_WEIGHT_CACHE.get(id(conv3d))
    if hit is not None and hit[0] == key:
        return hit[1]
    c1, b1, c2, b2, c3, b3, c11, b11 = mods
    if (tuple(c1.weight.shape) != (16, 1, 8, 3, 3) or tuple(c2.weight.shape) != (32, 16, 4, 3, 3)
            or tuple(c3.weight.shape) != (16, 32, 2, 3, 3) or tuple(c11.weight.shape) != (1, 16, 1, 1)):
        raise RuntimeError("v4_cost_volume: unexpected Conv3d / volume11 shapes (not MobileStereoNetV4's)")
    with torch.no_grad():
        w1, t1 = _fold_bn(c1.weight[:, 0], c1.bias, b1)                       # (16,8,3,3)
        w2, t2 = _fold_bn(c2.weight, c2.bias, b2)                             # (32,16,4,3,3) [co,ci,kd,dy,dx]
        w3, t3 = _fold_bn(c3.weight, c3.bias, b3)                             # (16,32,2,3,3)
        w11, t11 = _fold_bn(c11.weight, c11.bias, b11)                        # (1,16,1,1)

        def pack(w):                                                          # -> (9, 8, co, 8): kk = kd*Cin + ci
            co = w.shape[0]
            k = w.permute(3, 4, 2, 1, 0).reshape(9, 64, co)                   # [tap][kd*Cin + ci][co]
            return k.reshape(9, 8, 8, co).permute(0, 1, 3, 2).contiguous().to(op_dtype)

        packed = {"w1": w1.contiguous(), "t1": t1.contiguous(), "w2": pack(w2), "t2": t2.contiguous(), "w3": pack(w3),
                  "t3": t3.contiguous(), "w11": w11.reshape(16).contiguous(), "t11": t11.reshape(1).contiguous()}
    _V4_WEIGHT_CACHE[id(conv3d)] = (key, packed)
    return packed


def v4_cost_volume(left, right, conv3d, volume11, max_disparity, op_dtype=None):
    """MobileStereoNetV4's per-disparity learned volume (mobile_stereo_net_v4.py:443-458), all ``max_disparity``
    iterations and the three Conv3d + 1x1 conv in one fused op (eval mode: BatchNorm folded).  (B,32,H,W) x2 ->
    (B,D,H,W), zero where x < d.  Operands are fp16 (bf16 for bf16 features), accumulation fp32.  Inference only."""
    dev, n, c, h, w = _check_pair(left, right)
    if conv3d.training or volume11.training:
        raise RuntimeError("v4_cost_volume folds BatchNorm: eval mode only (the training forward keeps the loop)")
    d = int(max_disparity)
    op_dtype = op_dtype or (torch.bfloat16 if left.dtype == torch.bfloat16 else torch.float16)
    pk = pack_v4_weights(conv3d, volume11, op_dtype)
    lib = L.load()
    out = torch.empty((n, d, h, w), dtype=left.dtype, device=left.device)
    nbytes = lib.rsm_v4_volume_workspace(n, h, w, d)
    work = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=left.device)
    wts = L.RsmV4Weights(*(pk[k].data_ptr() for k in ("w1", "t1", "w2", "t2", "w3", "t3", "w11", "t11")))
    L.check(lib.rsm_v4_volume_fwd(L.feat(left.detach()), L.feat(right.detach()), wts, out.data_ptr(), work.data_ptr(),
                                  n, c, h, w, d, L.dtype_code(left), L._DTYPES[op_dtype], dev, L.stream_ptr(dev)),
            "rsm_v4_volume_fwd")
    return out


# ----------------------------------------------------------------------- refinement warp
class _Warp(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, image, flow):
        if image.dim() != 4 or flow.dim() != 4:
            raise ValueError(f"expected (N,C,H,W) image and (N,1|2,H,W) flow, got {tuple(image.shape)} and {tuple(flow.shape)}")
        n, c, h, w = image.shape
        cf = flow.shape[1]
        if (flow.shape[0], flow.shape[2], flow.shape[3]) != (n, h, w):
            raise RuntimeError(f"image {tuple(image.shape)} and flow {tuple(flow.shape)} shapes differ")
        if image.dtype != flow.dtype:
            raise TypeError(f"image ({image.dtype}) and flow ({flow.dtype}) dtypes differ")
        dev = L.require_cuda(image, flow)
        image, flow = _dense(image), _dense(flow)
        out = torch.empty_like(image)
        L.check(L.load().rsm_warp_fwd(image.data_ptr(), flow.data_ptr(), out.data_ptr(), n, c, h, w, cf,
                                      L.dtype_code(image), dev, L.stream_ptr(dev)), "rsm_warp_fwd")
        ctx.save_for_backward(image, flow)
        ctx.dims = (dev, n, c, h, w, cf)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        image, flow = ctx.saved_tensors
        dev, n, c, h, w, cf = ctx.dims
        gout = _dense(gout)
        # fp32 atomics accumulate the image gradient; skipped entirely when the image needs none (v2's RefineNet
        # warps the raw right image and differentiates only the flow)
        gimage = torch.empty((n, c, h, w), dtype=torch.float32, device=image.device) if ctx.needs_input_grad[0] else None
        gflow = torch.empty_like(flow) if ctx.needs_input_grad[1] else None
        if gimage is None and gflow is None:
            return None, None
        L.check(L.load().rsm_warp_bwd(gout.data_ptr(), image.data_ptr(), flow.data_ptr(), L.ptr(gimage), L.ptr(gflow),
                                      n, c, h, w, cf, L.dtype_code(image), dev, L.stream_ptr(dev)), "rsm_warp_bwd")
        return (None if gimage is None else gimage.to(image.dtype)), gflow


def warp_by_flow_map(image, flow):
    """warp_by_flow_map, model/mobile_stereo_net_v2.py:59-96 (= v3 :60-97, tools/warp.py:5-42): bilinear,
    zero-padded sample of ``image`` (N,C,H,W) at the pixel grid minus ``flow`` (N,1|2,H,W), with the reference's
    own coordinate convention ((size-1) normalisation seen by grid_sample(align_corners=False))."""
    cf = flow.shape[1] if flow.dim() == 4 else -1
    assert cf == 1 or cf == 2, f"invalid flow map dimension 1 or 2 ({cf})!"
    return _Warp.apply(image, flow)


# ---------------------------------------------------------------------------- pre / post steps
class _Prepare(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, img, hp, wp):
        dev = L.require_cuda(img)
        img = _dense(img)
        n, c, h, w = img.shape
        out = torch.empty((n, c, hp, wp), dtype=img.dtype, device=img.device)
        L.check(L.load().rsm_prepare_fwd(img.data_ptr(), out.data_ptr(), n * c, h, w, hp, wp, L.dtype_code(img), dev,
                                         L.stream_ptr(dev)), "rsm_prepare_fwd")
        ctx.dims = (dev, n, c, h, w, hp, wp)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, h, w, hp, wp = ctx.dims
        gout = _dense(gout)
        gimg = torch.empty((n, c, h, w), dtype=gout.dtype, device=gout.device)
        L.check(L.load().rsm_prepare_bwd(gout.data_ptr(), gimg.data_ptr(), n * c, h, w, hp, wp, L.dtype_code(gout), dev,
                                         L.stream_ptr(dev)), "rsm_prepare_bwd")
        return gimg, None, None


def prepare_input(img, align=1):
    """The models' first lines, model/mobile_stereo_net.py:121-130 (= _v2.py:194-203, _v3.py:296-305,
    mobile_disp_net_c.py:339-351; _v4.py:433-434 with ``align=1``): ``2 * (img / 255) - 1`` then zero padding on the
    right / bottom up to a multiple of ``align``.  (N,C,H,W) -> (N,C,Hp,Wp)."""
    if img.dim() != 4:
        raise ValueError(f"expected a (N,C,H,W) image, got {tuple(img.shape)}")
    if int(align) < 1:
        raise ValueError(f"align must be >= 1, got {align}")
    h, w = img.shape[2:]
    hp = h + (align - (h % align)) % align
    wp = w + (align - (w % align)) % align
    return _Prepare.apply(img, int(hp), int(wp))


_RESIZE_MODES = {"nearest": 0, "bilinear": 1}


class _Finalize(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, disp, hp, wp, h, w, vscale, mode):
        dev = L.require_cuda(disp)
        disp = _dense(disp)
        n, c, hs, ws = disp.shape
        out = torch.empty((n, c, h, w), dtype=disp.dtype, device=disp.device)
        L.check(L.load().rsm_finalize_fwd(disp.data_ptr(), out.data_ptr(), n * c, hs, ws, hp, wp, h, w, vscale, mode,
                                          L.dtype_code(disp), dev, L.stream_ptr(dev)), "rsm_finalize_fwd")
        ctx.dims = (dev, n, c, hs, ws, hp, wp, h, w, vscale, mode)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, c, hs, ws, hp, wp, h, w, vscale, mode = ctx.dims
        gout = _dense(gout)
        gdisp = torch.empty((n, c, hs, ws), dtype=gout.dtype, device=gout.device)
        L.check(L.load().rsm_finalize_bwd(gout.data_ptr(), gdisp.data_ptr(), n * c, hs, ws, hp, wp, h, w, vscale, mode,
                                          L.dtype_code(gout), dev, L.stream_ptr(dev)), "rsm_finalize_bwd")
        return gdisp, None, None, None, None, None, None


def finalize_disparity(disp, padded_size, size=None, mode="nearest", negate=True):
    """The models' last lines: ``-1.0 * F.interpolate(disp * scale, padded_size)[:, :, :h, :w]`` with
    ``scale = padded_W / disp_W`` -- model/mobile_stereo_net.py:154-159 (= _v2.py:227-232; ``mode="nearest"``, the
    F.interpolate default) and ``disparity_interpolate`` + crop + negate, model/mobile_disp_net_c.py:223-234 +
    :408-411 (``mode="bilinear"``, align_corners=False; a same-size map is only cropped and negated there).
    (N,C,hs,ws) -> (N,C,h,w); ``size`` defaults to ``padded_size``; ``negate=False`` keeps the sign (the resize
    alone, e.g. as a drop-in for ``disparity_interpolate``)."""
    if disp.dim() != 4:
        raise ValueError(f"expected a (N,C,h,w) disparity map, got {tuple(disp.shape)}")
    if mode not in _RESIZE_MODES:
        raise ValueError(f"mode must be 'nearest' or 'bilinear', got {mode!r}")
    hp, wp = (int(v) for v in padded_size)
    h, w = (hp, wp) if size is None else (int(v) for v in size)
    if not (0 <= h <= hp and 0 <= w <= wp):
        raise ValueError(f"crop size {(h, w)} exceeds the resized map {(hp, wp)}")
    hs, ws = disp.shape[2:]
    vscale = float(wp) / ws if ws > 0 else 1.0
    if mode == "bilinear" and (hs, ws) == (hp, wp):
        vscale = 1.0      # disparity_interpolate leaves a same-size map untouched (mobile_disp_net_c.py:228)
    # every rounding in the kernel is sign-symmetric, so "do not negate" is exactly a negated value scale
    return _Finalize.apply(disp, hp, wp, h, w, vscale if negate else -vscale, _RESIZE_MODES[mode])


# ---------------------------------------------------------------------------- loss / metrics
def _reduce_workspace(device):
    return torch.empty((L.RSM_REDUCE_WS_DOUBLES,), dtype=torch.float64, device=device)


class _SeqLossTerm(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, pred, gt, valid, max_flow, kind):
        dev = L.require_cuda(pred, gt, valid)
        pred, gt, valid = _dense(pred), _dense(gt), _dense(valid)
        n, _, hs, ws = pred.shape
        h, w = gt.shape[2:]
        result = torch.empty((4,), dtype=torch.float64, device=pred.device)
        L.check(L.load().rsm_seqloss_fwd(pred.data_ptr(), gt.data_ptr(), valid.data_ptr(), _reduce_workspace(pred.device).data_ptr(),
                                         result.data_ptr(), n, hs, ws, h, w, max_flow, kind, L.dtype_code(pred), dev,
                                         L.stream_ptr(dev)), "rsm_seqloss_fwd")
        ctx.save_for_backward(pred, gt, valid, result)
        ctx.dims = (dev, n, hs, ws, h, w, max_flow, kind)
        ctx.mark_non_differentiable(result)
        return result[0].to(torch.float32), result

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gmean, _gresult):
        pred, gt, valid, result = ctx.saved_tensors
        dev, n, hs, ws, h, w, max_flow, kind = ctx.dims
        gmean = gmean.to(torch.float32).contiguous()
        gpred = torch.empty_like(pred)
        L.check(L.load().rsm_seqloss_bwd(gmean.data_ptr(), result.data_ptr(), pred.data_ptr(), gt.data_ptr(), valid.data_ptr(),
                                         gpred.data_ptr(), n, hs, ws, h, w, max_flow, kind, L.dtype_code(pred), dev,
                                         L.stream_ptr(dev)), "rsm_seqloss_bwd")
        return gpred, None, None, None, None


def sequence_loss_term(pred, flow_gt, flow_valid, max_flow=700.0, smooth=False):
    """One term of SequenceLoss.forward (loss/loss.py:55-80): the masked mean of |gt - p| (``smooth``: smooth-L1,
    beta 1) with p = F.interpolate(pred * scale, gt size) when the sizes differ.  Returns ``(mean, stats)``:
    a differentiable fp32 scalar and the fp64 vector {mean, valid count, non-finite predictions, infinite valid gt}."""
    if pred.dim() != 4 or flow_gt.dim() != 4 or pred.shape[1] != 1 or flow_gt.shape[1] != 1:
        raise ValueError(f"expected (N,1,h,w) prediction and (N,1,H,W) ground truth, got {tuple(pred.shape)} and {tuple(flow_gt.shape)}")
    if flow_valid.shape != (flow_gt.shape[0],) + tuple(flow_gt.shape[2:]) or pred.shape[0] != flow_gt.shape[0]:
        raise RuntimeError(f"shapes differ: pred {tuple(pred.shape)}, gt {tuple(flow_gt.shape)}, valid {tuple(flow_valid.shape)}")
    if pred.dtype != flow_gt.dtype:                 # torch's type promotion in `gt - pred`
        dt = torch.promote_types(pred.dtype, flow_gt.dtype)
        pred, flow_gt = pred.to(dt), flow_gt.to(dt)
    return _SeqLossTerm.apply(pred, flow_gt, flow_valid.to(torch.float32), float(max_flow), 1 if smooth else 0)


def flow_map_metrics(flow_gt, flow_pred, flow_valid):
    """get_flow_map_metrics (loss/loss.py:6-22) as ONE pass; returns the device vector (fp64)
    {epe, 0.5px, 1px, 3px, 5px, min, max, valid count} without synchronising."""
    if flow_gt.dim() != 4 or flow_gt.shape != flow_pred.shape:
        raise RuntimeError(f"flow_gt {tuple(flow_gt.shape)} and flow_pred {tuple(flow_pred.shape)} shapes differ")
    n, c, h, w = flow_gt.shape
    if tuple(flow_valid.shape) != (n, h, w):
        raise RuntimeError(f"flow_valid {tuple(flow_valid.shape)} does not match {(n, h, w)}")
    if flow_gt.dtype != flow_pred.dtype:
        raise TypeError(f"flow_gt ({flow_gt.dtype}) and flow_pred ({flow_pred.dtype}) dtypes differ")
    dev = L.require_cuda(flow_gt, flow_pred, flow_valid)
    gt, pred, valid = _dense(flow_gt.detach()), _dense(flow_pred.detach()), _dense(flow_valid.detach().to(torch.float32))
    result = torch.empty((8,), dtype=torch.float64, device=gt.device)
    L.check(L.load().rsm_flow_metrics(gt.data_ptr(), pred.data_ptr(), valid.data_ptr(), _reduce_workspace(gt.device).data_ptr(),
                                      result.data_ptr(), n, c, h, w, L.dtype_code(gt), dev, L.stream_ptr(dev)), "rsm_flow_metrics")
    return result


# ---------------------------------------------------------------------------- regression
def _regress_outputs(shape, dtype, device, soft, argmin, argmax, lse, expect=False):
    """Output planes of a regression call.  ``expect`` is the fp32 copy of the expectation (rsm_regress_out.expect):
    the backward pass reads it, and it is what a caller under autocast returns; for an fp32 cost it IS ``soft``."""
    n, h, w = shape
    so = torch.empty((n, h, w), dtype=dtype, device=device) if soft else None
    mi = torch.empty((n, h, w), dtype=torch.int64, device=device) if argmin else None
    ma = torch.empty((n, h, w), dtype=torch.int64, device=device) if argmax else None
    ls = torch.empty((n, h, w), dtype=torch.float32, device=device) if lse else None
    ex = None
    if expect:
        ex = so if (so is not None and dtype == torch.float32) else torch.empty((n, h, w), dtype=torch.float32, device=device)
    return so, mi, ma, ls, ex, L.RsmRegressOut(L.ptr(so), L.ptr(mi), L.ptr(ma), L.ptr(ls), L.ptr(ex))


def _check_cost(cost):
    if cost.dim() != 4:
        raise ValueError(f"expected a (N,D,H,W) cost volume, got {tuple(cost.shape)}")
    if cost.shape[1] == 0:
        raise ValueError("cost volume has an empty disparity axis")
    return L.require_cuda(cost)


class _Regress(torch.autograd.Function):
    """soft-argmax + hard argmin + hard argmax in ONE pass over the volume."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, cost, want_argmin, want_argmax, out_fp32=False):
        dev = _check_cost(cost)
        cost = _dense(cost)
        n, d, h, w = cost.shape
        need_grad = ctx.needs_input_grad[0]
        out_fp32 = bool(out_fp32) and cost.dtype != torch.float32
        so, mi, ma, ls, ex, out = _regress_outputs((n, h, w), cost.dtype, cost.device, not out_fp32, want_argmin,
                                                   want_argmax, need_grad, need_grad or out_fp32)
        L.check(L.load().rsm_regress_fwd(cost.data_ptr(), n, d, h, w, L.dtype_code(cost), out, dev,
                                         L.stream_ptr(dev)), "rsm_regress_fwd")
        if out_fp32:
            so = ex
        if need_grad:
            ctx.save_for_backward(cost, ex, ls)
        ctx.dims = (dev, n, d, h, w)
        empty = torch.empty(0, dtype=torch.int64, device=cost.device)
        mi = mi if mi is not None else empty
        ma = ma if ma is not None else empty
        ctx.mark_non_differentiable(mi, ma)
        return so, mi, ma

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gsoft, _gmi, _gma):
        cost, ex, ls = ctx.saved_tensors
        dev, n, d, h, w = ctx.dims
        gsoft = _dense(gsoft.to(cost.dtype))
        gcost = torch.empty_like(cost)
        L.check(L.load().rsm_regress_bwd(gsoft.data_ptr(), cost.data_ptr(), ex.data_ptr(), ls.data_ptr(),
                                         gcost.data_ptr(), n, d, h, w, L.dtype_code(cost), dev,
                                         L.stream_ptr(dev)), "rsm_regress_bwd")
        return gcost, None, None, None


def regress(cost, argmin=True, argmax=True):
    """One pass over a (N,D,H,W) cost -> (soft (N,H,W), argmin (N,H,W) int64, argmax (N,H,W) int64).

    soft = sum_d d * softmax_d(+cost) (reference mobile_stereo_net.py:144-147); argmin/argmax follow
    torch.argmin/argmax(cost, 1): first index on ties, NaN is the extremum."""
    so, mi, ma = _Regress.apply(cost, bool(argmin), bool(argmax), False)
    return so, (mi if argmin else None), (ma if argmax else None)


def soft_argmax(cost, keepdim=False, out_fp32=False):
    """sum_d d * softmax_d(+cost): (N,D,H,W) -> (N,H,W), or (N,1,H,W) with ``keepdim``.  ``out_fp32``: return the
    fp32 expectation for a 16-bit cost (what the reference produces under autocast, where F.softmax runs in fp32)."""
    so, _, _ = _Regress.apply(cost, False, False, out_fp32)
    return so.unsqueeze(1) if keepdim else so


def hard_argmin(cost):
    """torch.argmin(cost, dim=1) for a (N,D,H,W) cost, int64."""
    dev = _check_cost(cost)
    cost = _dense(cost.detach())
    n, d, h, w = cost.shape
    _, mi, _, _, _, out = _regress_outputs((n, h, w), cost.dtype, cost.device, False, True, False, False)
    L.check(L.load().rsm_regress_fwd(cost.data_ptr(), n, d, h, w, L.dtype_code(cost), out, dev,
                                     L.stream_ptr(dev)), "rsm_regress_fwd")
    return mi


def hard_argmax(cost):
    """torch.argmax(cost, dim=1) for a (N,D,H,W) cost, int64."""
    dev = _check_cost(cost)
    cost = _dense(cost.detach())
    n, d, h, w = cost.shape
    _, _, ma, _, _, out = _regress_outputs((n, h, w), cost.dtype, cost.device, False, False, True, False)
    L.check(L.load().rsm_regress_fwd(cost.data_ptr(), n, d, h, w, L.dtype_code(cost), out, dev,
                                     L.stream_ptr(dev)), "rsm_regress_fwd")
    return ma


class _Expect(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, prob):
        dev = _check_cost(prob)
        prob = _dense(prob)
        n, d, h, w = prob.shape
        out = torch.empty((n, h, w), dtype=prob.dtype, device=prob.device)
        L.check(L.load().rsm_expect_fwd(prob.data_ptr(), out.data_ptr(), n, d, h, w, L.dtype_code(prob), dev,
                                        L.stream_ptr(dev)), "rsm_expect_fwd")
        ctx.dims = (dev, n, d, h, w)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        dev, n, d, h, w = ctx.dims
        gout = _dense(gout)
        gp = torch.empty((n, d, h, w), dtype=gout.dtype, device=gout.device)
        L.check(L.load().rsm_expect_bwd(gout.data_ptr(), gp.data_ptr(), n, d, h, w, L.dtype_code(gout), dev,
                                        L.stream_ptr(dev)), "rsm_expect_bwd")
        return gp


def expectation(prob):
    """sum_d d * prob[:, d] for already-normalised probabilities (mobile_stereo_net_v4.py:10-14)."""
    return _Expect.apply(prob)


# -------------------------------------------------------------------------- v4 head
class _UpsampleRegress(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, cost, maxdisp, out_h, out_w, want_argmin, want_argmax, out_fp32=False):
        dev = _check_cost(cost)
        cost = _dense(cost)
        b, dc, hc, wc = cost.shape
        d, h, w = int(maxdisp), int(out_h), int(out_w)
        need_grad = ctx.needs_input_grad[0]
        out_fp32 = bool(out_fp32) and cost.dtype != torch.float32
        so, mi, ma, ls, ex, out = _regress_outputs((b, h, w), cost.dtype, cost.device, not out_fp32, want_argmin,
                                                   want_argmax, need_grad, need_grad or out_fp32)
        L.check(L.load().rsm_upsample_regress_fwd(cost.data_ptr(), b, dc, hc, wc, d, h, w, L.dtype_code(cost),
                                                  out, dev, L.stream_ptr(dev)), "rsm_upsample_regress_fwd")
        if out_fp32:
            so = ex
        if need_grad:
            ctx.save_for_backward(cost, ex, ls)
        ctx.dims = (dev, b, dc, hc, wc, d, h, w)
        empty = torch.empty(0, dtype=torch.int64, device=cost.device)
        mi = mi if mi is not None else empty
        ma = ma if ma is not None else empty
        ctx.mark_non_differentiable(mi, ma)
        return so, mi, ma

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gsoft, _gmi, _gma):
        cost, ex, ls = ctx.saved_tensors
        dev, b, dc, hc, wc, d, h, w = ctx.dims
        gsoft = _dense(gsoft.to(cost.dtype))
        gcost = torch.empty_like(cost)
        nbytes = L.load().rsm_upsample_regress_bwd_workspace(b, dc, h, w)
        work = torch.empty(nbytes // 4, dtype=torch.float32, device=cost.device)
        L.check(L.load().rsm_upsample_regress_bwd(gsoft.data_ptr(), cost.data_ptr(), ex.data_ptr(),
                                                  ls.data_ptr(), gcost.data_ptr(), work.data_ptr(), b, dc, hc,
                                                  wc, d, h, w, L.dtype_code(cost), dev, L.stream_ptr(dev)),
                "rsm_upsample_regress_bwd")
        return gcost, None, None, None, None, None, None


def upsample_regress(cost, maxdisp, out_h, out_w, argmin=False, argmax=False, out_fp32=False):
    """MobileStereoNetV4 head (mobile_stereo_net_v4.py:511-518), fused: trilinear upsample of the coarse
    (B,Dc,Hc,Wc) cost to (maxdisp,out_h,out_w) -> softmax over D -> expectation, (B,H,W).
    Optionally also the hard argmin/argmax over the upsampled volume."""
    so, mi, ma = _UpsampleRegress.apply(cost, maxdisp, out_h, out_w, bool(argmin), bool(argmax), bool(out_fp32))
    if not argmin and not argmax:
        return so
    return so, (mi if argmin else None), (ma if argmax else None)


# --------------------------------------------------------------- fused build + regress
def inner_product_regress(left, right, max_disparity, mean=False, argmin=True, argmax=True):
    """Inner-product / correlation volume reduced on chip to (soft fp32, argmin, argmax); the
    (N,D,H,W) volume is never written to HBM.  Inference only (no autograd)."""
    dev, n, c, h, w = _check_pair(left, right)
    d = int(max_disparity)
    so, mi, ma, _, _, out = _regress_outputs((n, h, w), torch.float32, left.device, True, argmin, argmax, False)
    red = L.RSM_REDUCE_MEAN if mean else L.RSM_REDUCE_SUM
    L.check(L.load().rsm_inner_regress_fwd(L.feat(left.detach()), L.feat(right.detach()), n, c, h, w, d, red,
                                           L.dtype_code(left), out, dev, L.stream_ptr(dev)),
            "rsm_inner_regress_fwd")
    return so, mi, ma
