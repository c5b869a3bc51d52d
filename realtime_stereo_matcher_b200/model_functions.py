"""Model-local hot-path functions of the reference under their own names and signatures
(SURVEY.md 8b).  The reference models look these up as module globals at call time, so
``patch.patch_reference()`` can swap them in without touching weights or state_dict keys.
"""
from __future__ import annotations

import torch

from . import functional as F_rsm


def _autocast_fp32() -> bool:
    """Under torch.autocast the reference's F.softmax runs in fp32 (autocast's fp32 op list), so its regressed
    disparity is fp32 even when the cost is fp16 (evaluate_stereo.py / test_stereo.py wrap the
    forward in autocast).  The kernels then return the fp32 expectation plane instead of rounding it to 16 bit."""
    return torch.is_autocast_enabled("cuda")


def make_cost_volume(left, right, max_disp):
    """model/mobile_stereo_net.py:8-27 (= mobile_stereo_net_v2.py:8-27, mobile_stereo_net_v3.py:9-28):
    (N,C,H,W) x2 -> (N,C,max_disp,H,W) difference volume, out-of-range region filled with 1.0."""
    return F_rsm.difference_volume(left, right, max_disp, fill=1.0)


def make_correlation_volume(l_fmap, r_fmap, max_disp):
    """model/mobile_disp_net_c.py:188-205: (N,C,H,W) x2 -> (N,max_disp,H,W) channel-MEAN correlation."""
    return F_rsm.inner_product_volume(l_fmap, r_fmap, max_disp, mean=True)


def interweave_tensors(refimg_fea, targetimg_fea):
    """model/mobile_stereo_net_v4.py:17-23: (B,C,H,W) x2 -> (B,2C,H,W), even = ref, odd = target.
    Accepts the width-cropped, non-contiguous slices the v4 forward passes (:446)."""
    return F_rsm.interweave(refimg_fea, targetimg_fea)


def shift_interweave_stack(refimg_fea, targetimg_fea, volume_size):
    """All ``volume_size`` inputs of the MobileStereoNetV4 per-disparity loop at once
    (model/mobile_stereo_net_v4.py:444-458): (B,C,H,W) x2 -> (D,B,2C,H,W) with
    out[i][..., i:] == interweave_tensors(ref[..., i:], target[..., :-i]) and zeros for x < i."""
    return F_rsm.shift_interweave_volume(refimg_fea, targetimg_fea, volume_size)


def v4_cost_volume(featL, featR, conv3d, volume11, volume_size):
    """The whole per-disparity volume loop of MobileStereoNetV4.forward (model/mobile_stereo_net_v4.py:443-458) with
    the module's own ``conv3d`` / ``volume11`` weights, eval mode: (B,32,H,W) x2 -> (B,volume_size,H,W)."""
    return F_rsm.v4_cost_volume(featL, featR, conv3d, volume11, volume_size)


def warp_by_flow_map(image, flow):
    """model/mobile_stereo_net_v2.py:59-96 (= mobile_stereo_net_v3.py:60-97, tools/warp.py:5-42): the
    refinement warp of RefineNet (call sites v2 :127, v3 :136); same AssertionError on a bad flow shape."""
    if image.dtype != flow.dtype or (_autocast_fp32() and image.dtype != torch.float32):
        # the reference's F.grid_sample is on autocast's fp32 list (and type-promotes otherwise): under autocast the
        # v3 RefineNet warps fp16 feature maps with the fp32 disparity and gets an fp32 map back
        dt = torch.float32 if _autocast_fp32() else torch.promote_types(image.dtype, flow.dtype)
        image, flow = image.to(dt), flow.to(dt)
    return F_rsm.warp_by_flow_map(image, flow)


def prepare_input(img, align=1):
    """model/mobile_stereo_net.py:121-130 (= _v2.py:194-203, _v3.py:296-305, mobile_disp_net_c.py:339-351,
    _v4.py:433-434 with align=1): 2 * (img / 255) - 1, zero-padded right / bottom to a multiple of ``align``."""
    return F_rsm.prepare_input(img, align)


def finalize_disparity(x, padded_size, size):
    """model/mobile_stereo_net.py:156 + :159 (= _v2.py:229 + :232, _v3.py): -1.0 * F.interpolate(x * scale,
    padded_size)[:, :, :h, :w] with scale = padded_W / x_W."""
    return F_rsm.finalize_disparity(x, padded_size, size, mode="nearest")


def disparity_interpolate(disp, shape):
    """model/mobile_disp_net_c.py:223-234: bilinear (align_corners=False) resize of ``disp * (dst_w / src_w)``;
    a map already at ``shape`` is returned untouched."""
    if tuple(disp.shape[2:]) == tuple(shape):
        return disp
    if _autocast_fp32() and disp.dtype != torch.float32:
        disp = disp.float()      # F.interpolate (upsample_bilinear2d) is on autocast's fp32 list: fp32 in, fp32 out
    return F_rsm.finalize_disparity(disp, shape, None, mode="bilinear", negate=False)


def disparity_regression_v4(x, maxdisp):
    """model/mobile_stereo_net_v4.py:10-14: x holds PROBABILITIES (already softmax-ed);
    returns sum_d d * x[:, d] as (N,H,W)."""
    assert len(x.shape) == 4
    assert x.shape[1] == maxdisp, f"disparity axis ({x.shape[1]}) != maxdisp ({maxdisp})"
    return F_rsm.expectation(x)


def disparity_regression_dispnetc(corr_volume, max_disp):
    """model/mobile_disp_net_c.py:208-220: corr_volume holds LOGITS; softmax over dim 1 then
    expectation, returned as (N,1,H,W)."""
    assert len(corr_volume.shape) == 4, "#dimensions of correlation volume != 4."
    assert corr_volume.shape[1] == max_disp, f"#channels of correlation volume != max_disparity ({max_disp})."
    return F_rsm.soft_argmax(corr_volume, keepdim=True, out_fp32=_autocast_fp32())


def softmax_regression(cost_volume, keepdim=True):
    """The inline regression of MobileStereoNet v1-v3 (mobile_stereo_net.py:144-147,
    mobile_stereo_net_v2.py:217-220, mobile_stereo_net_v3.py:321-324): softmax(dim=1) ->
    sum(x * arange(D)), one fused pass."""
    return F_rsm.soft_argmax(cost_volume, keepdim=keepdim, out_fp32=_autocast_fp32())


def v4_head(cost, maxdisp, out_h, out_w):
    """The MobileStereoNetV4 head (mobile_stereo_net_v4.py:511-518; training heads :471-506):
    F.interpolate(cost[:,None], [maxdisp,H,W], 'trilinear') -> softmax -> disparity_regression,
    without materialising the (B,maxdisp,H,W) tensor.  Returns (B,H,W)."""
    return F_rsm.upsample_regress(cost, maxdisp, out_h, out_w, out_fp32=_autocast_fp32())
