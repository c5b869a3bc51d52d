"""Swap the reference's hot path for the B200 kernels, in place, without touching weights.

    import realtime_stereo_matcher_b200 as rsm
    rsm.patch_reference()            # reference checkout must be importable (sys.path)
    net = model.build_model(cfg["model"]).cuda()     # the reference's own factory
    net(left, right)                 # cost volumes + regression now run in librsm_b200.so

Two levels:
  1. module-global functions the reference models resolve at call time
     (make_cost_volume, make_correlation_volume, interweave_tensors, disparity_regression) and the
     four ``cost_volume.*`` classes are replaced by the mirrors in this package;
  2. with ``fuse=True`` the model ``forward`` methods are replaced by the glue below, which calls
     the SAME sub-modules in the same order (state_dict keys and numerics of every conv stack are
     untouched) but routes the inline ``F.softmax -> arange -> sum`` regression of v1-v3 and the
     ``F.interpolate(trilinear) -> softmax -> disparity_regression`` head of v4 through the fused
     single-pass kernels.
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Tuple

from . import cost_volume as cv_mirror
from . import model_functions as mf

_SAVED: List[Tuple[object, str, object]] = []


def _set(obj, name, value):
    _SAVED.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


# ------------------------------------------------------------------ fused forward glue
def _normalise_and_pad(self, l_img, r_img):
    # mobile_stereo_net.py:121-130 (same in v2 :194-203, v3 :296-305)
    h, w = l_img.shape[2:]
    return mf.prepare_input(l_img, self.align), mf.prepare_input(r_img, self.align), h, w


def _refine_outputs(x, steps, l_img, h, w):
    # mobile_stereo_net.py:149-159: refine, rescale to full resolution, crop, negate
    outs = []
    for step in steps:
        x = step(x)
        outs.append(mf.finalize_disparity(x, l_img.shape[2:], (h, w)))
    return outs


def forward_v1(self, left_img, right_img):
    """MobileStereoNet.forward (mobile_stereo_net.py:120-159) with the fused hot path."""
    l_img, r_img, h, w = _normalise_and_pad(self, left_img, right_img)
    lf, rf = self.feature_extractor(l_img), self.feature_extractor(r_img)
    cost = self.cost_filter(mf.make_cost_volume(lf, rf, self.max_disp)).squeeze(1)
    x = mf.softmax_regression(cost, keepdim=True)
    return _refine_outputs(x, [lambda t, r=r: r(t, l_img) for r in self.refine_layer], l_img, h, w)


def forward_v2(self, left_img, right_img):
    """MobileStereoNetV2.forward (mobile_stereo_net_v2.py:193-232) with the fused hot path."""
    l_img, r_img, h, w = _normalise_and_pad(self, left_img, right_img)
    lf, rf = self.feature_extractor(l_img), self.feature_extractor(r_img)
    cost = self.cost_filter(mf.make_cost_volume(lf, rf, self.max_disp)).squeeze(1)
    x = mf.softmax_regression(cost, keepdim=True)
    return _refine_outputs(x, [lambda t, r=r: r(t, l_img, r_img) for r in self.refine_layer], l_img, h, w)


def forward_v3(self, l_img, r_img):
    """MobileStereoNetV3.forward (mobile_stereo_net_v3.py:295-336) with the fused hot path."""
    l_img, r_img, h, w = _normalise_and_pad(self, l_img, r_img)
    lfs, rfs = self.feature_extractor(l_img), self.feature_extractor(r_img)
    cost = self.cost_filter(mf.make_cost_volume(lfs[0], rfs[0], self.max_disp)).squeeze(1)
    x = mf.softmax_regression(cost, keepdim=True)
    steps = [lambda t, r=r, i=i: r(t, lfs[i + 1], rfs[i + 1]) for i, r in enumerate(self.refine_layers)]
    return _refine_outputs(x, steps, l_img, h, w)


def v4_volume_batched(self, featL, featR, chunk=None):
    """MobileStereoNetV4's per-disparity volume loop (mobile_stereo_net_v4.py:443-458) as ONE batched pass
    (eval mode: BatchNorm is affine, so batching the 48 iterations is exact).  The 48 interweaved, shifted
    inputs come from one kernel (rsm_shift_interweave_fwd) at full width with zeros where x < d; the
    reference's convolutions on the cropped tensors see zero padding at the crop edge, which is reproduced
    by zeroing the x < d columns again after every ReLU.  The convolutions stay the module's own (cuDNN)."""
    import torch
    B, C, H, W = featL.shape
    D = self.volume_size
    out = featL.new_zeros([B, D, H, W])
    chunk = chunk or D
    xs = torch.arange(W, device=featL.device)
    for d0 in range(0, D, chunk):                       # optional chunking over disparities bounds memory
        d1 = min(D, d0 + chunk)
        if d0 == 0 and d1 == D:
            x = mf.shift_interweave_stack(featL, featR, D)                       # (D,B,2C,H,W)
        else:
            x = mf.shift_interweave_stack(featL, featR, d1)[d0:d1]
        nd = d1 - d0
        keep = (xs[None, :] >= torch.arange(d0, d1, device=featL.device)[:, None]).to(featL.dtype)   # (nd,W)
        keep = keep.repeat_interleave(B, dim=0)                                                      # (nd*B,W)
        x = x.reshape(nd * B, 1, 2 * C, H, W)
        for layer in self.conv3d:
            x = layer(x)
            if isinstance(layer, torch.nn.ReLU):
                x = x * keep.view(nd * B, 1, 1, 1, W)
        x = self.volume11(x.squeeze(2)) * keep.view(nd * B, 1, 1, W)                                 # (nd*B,1,H,W)
        out[:, d0:d1] = x.view(nd, B, H, W).permute(1, 0, 2, 3)
    return out


def forward_v4(self, L, R):
    import torch
    """MobileStereoNetV4.forward (mobile_stereo_net_v4.py:432-524) with interweave and the
    trilinear -> softmax -> expectation head on the fused kernels."""
    L, R = mf.prepare_input(L), mf.prepare_input(R)          # :433-434 (v4 does not pad)
    featL = self.preconv11(self.feature_extraction(L))
    featR = self.preconv11(self.feature_extraction(R))
    B, C, H, W = featL.shape
    if self.training:
        # BatchNorm3d normalises with the statistics of each iteration's own (cropped) batch, so the
        # per-disparity loop (:444-458) is kept as it is, with the interweave on the kernels
        volume = featL.new_zeros([B, self.num_groups, self.volume_size, H, W])
        for i in range(self.volume_size):
            x = mf.interweave_tensors(featL[:, :, :, i:], featR[:, :, :, : W - i])
            x = self.volume11(self.conv3d(x.unsqueeze(1)).squeeze(2))
            volume[:, :, i, :, i:] = x
        volume = volume.squeeze(1)
    elif torch.is_grad_enabled() and (featL.requires_grad or featR.requires_grad):
        volume = v4_volume_batched(self, featL, featR)         # eval mode but differentiable: cuDNN convolutions
    else:
        # the whole loop -- 48 x (interweave -> 3 Conv3d -> 1x1 conv) -- as one fused op on tcgen05
        volume = mf.v4_cost_volume(featL, featR, self.conv3d, self.volume11, self.volume_size)
    cost0 = self.dres0(volume)
    cost0 = self.dres1(cost0) + cost0
    out1 = self.encoder_decoder1(cost0)
    out2 = self.encoder_decoder2(out1)
    out3 = self.encoder_decoder3(out2)
    h, w = L.shape[2:]
    if self.training:
        costs = [self.classif0(cost0), self.classif1(out1), self.classif2(out2), self.classif3(out3)]
    else:
        costs = [self.classif3(out3)]
    return [-1.0 * mf.v4_head(c, self.maxdisp, h, w).unsqueeze(1) for c in costs]


# ---------------------------------------------------------------------------- patching
def patch_reference(fuse: bool = True, model_package: str = "model", cost_volume_package: str = "cost_volume",
                    loss_package: str = "loss") -> Dict[str, List[str]]:
    """Patch an importable reference checkout in place.  Returns what was patched.

    Raises ImportError when the reference packages cannot be imported (nothing is patched)."""
    done: Dict[str, List[str]] = {"functions": [], "classes": [], "forwards": [], "loss": []}
    m1 = importlib.import_module(f"{model_package}.mobile_stereo_net")
    m2 = importlib.import_module(f"{model_package}.mobile_stereo_net_v2")
    m3 = importlib.import_module(f"{model_package}.mobile_stereo_net_v3")
    m4 = importlib.import_module(f"{model_package}.mobile_stereo_net_v4")
    mc = importlib.import_module(f"{model_package}.mobile_disp_net_c")
    for m in (m1, m2, m3):
        _set(m, "make_cost_volume", mf.make_cost_volume)
        done["functions"].append(f"{m.__name__}.make_cost_volume")
    _set(mc, "make_correlation_volume", mf.make_correlation_volume)
    _set(mc, "disparity_regression", mf.disparity_regression_dispnetc)
    _set(mc, "disparity_interpolate", mf.disparity_interpolate)
    _set(m4, "interweave_tensors", mf.interweave_tensors)
    _set(m4, "disparity_regression", mf.disparity_regression_v4)
    done["functions"] += [f"{mc.__name__}.make_correlation_volume", f"{mc.__name__}.disparity_regression",
                          f"{mc.__name__}.disparity_interpolate",
                          f"{m4.__name__}.interweave_tensors", f"{m4.__name__}.disparity_regression"]
    for m in (m2, m3):   # refinement warp
        _set(m, "warp_by_flow_map", mf.warp_by_flow_map)
        done["functions"].append(f"{m.__name__}.warp_by_flow_map")
    try:
        for sub, cls in (("concatenate", "TorchConcatenateCost"), ("interweave", "TorchInterweaveCost"),
                         ("inner_product", "TorchInnerProductCost"), ("groupwise", "TorchGroupwiseCost")):
            mod = importlib.import_module(f"{cost_volume_package}.{sub}")
            _set(mod, cls, getattr(cv_mirror, cls))
            done["classes"].append(f"{mod.__name__}.{cls}")
    except ImportError:
        pass  # cost_volume/ is imported by nothing in the reference: optional
    try:   # loss / metrics on the device; scripts that did `from loss.loss import ...` before
        # this call keep their own binding -- patch first, or use the import swap of INTEGRATION.md
        from . import loss as loss_mirror
        lpkg = importlib.import_module(loss_package)
        lmod = importlib.import_module(f"{loss_package}.loss")
        _set(lpkg, "SequenceLoss", loss_mirror.SequenceLoss)
        _set(lmod, "SequenceLoss", loss_mirror.SequenceLoss)
        _set(lmod, "get_flow_map_metrics", loss_mirror.get_flow_map_metrics)
        done["loss"] += [f"{lpkg.__name__}.SequenceLoss", f"{lmod.__name__}.SequenceLoss", f"{lmod.__name__}.get_flow_map_metrics"]
    except ImportError:
        pass
    if fuse:
        for mod, cls, fwd in ((m1, "MobileStereoNet", forward_v1), (m2, "MobileStereoNetV2", forward_v2),
                              (m3, "MobileStereoNetV3", forward_v3), (m4, "MobileStereoNetV4", forward_v4)):
            _set(getattr(mod, cls), "forward", fwd)
            done["forwards"].append(f"{mod.__name__}.{cls}.forward")
    return done


def unpatch_reference() -> None:
    """Undo every patch_reference() call (restores the reference's own functions)."""
    while _SAVED:
        obj, name, value = _SAVED.pop()
        setattr(obj, name, value)
