"""CPU suite (build container only: needs /root/reference): the forward glue installed by
patch_reference(fuse=True) reproduces the reference models' own forwards.  The CUDA ops cannot run
here, so for THIS test they are backed by the numpy oracle (the oracle as checker of the glue);
the ops themselves are checked against the oracle on the GPU in tests/test_gpu_parity.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import oracle

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REF = os.environ.get("RSM_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")),
                                reason="reference checkout not present (GPU box)")


def _np(t):
    return t.detach().cpu().numpy()


def _t(a, like):
    return torch.from_numpy(np.ascontiguousarray(a)).to(like.dtype)


@pytest.fixture()
def oracle_backed(monkeypatch):
    """Route the functional ops through the oracle for the duration of one test."""
    from realtime_stereo_matcher_b200 import functional as F_rsm
    monkeypatch.setattr(F_rsm, "difference_volume", lambda l, r, d, fill=1.0: _t(oracle.difference_volume(_np(l), _np(r), d), l))
    monkeypatch.setattr(F_rsm, "inner_product_volume",
                        lambda l, r, d, mean=False, out_dtype=None: _t(oracle.inner_product_volume(_np(l), _np(r), d, mean=mean), l))
    monkeypatch.setattr(F_rsm, "interweave", lambda l, r: _t(oracle.interweave(_np(l), _np(r)), l))
    monkeypatch.setattr(F_rsm, "soft_argmax", lambda c, keepdim=False, out_fp32=False: _t(oracle.soft_argmax(_np(c), keepdim=keepdim), c))
    monkeypatch.setattr(F_rsm, "expectation",
                        lambda p: _t((_np(p) * np.arange(p.shape[1], dtype=np.float32).reshape(1, -1, 1, 1)).sum(1), p))
    monkeypatch.setattr(F_rsm, "upsample_regress", lambda c, d, h, w, argmin=False, argmax=False, out_fp32=False: _t(oracle.v4_tail(_np(c), d, h, w), c))
    monkeypatch.setattr(F_rsm, "prepare_input", lambda img, align=1: _t(oracle.prepare_input(_np(img), align), img))
    monkeypatch.setattr(F_rsm, "finalize_disparity",
                        lambda d, padded, size=None, mode="nearest", negate=True:
                        _t((1.0 if negate else -1.0) * oracle.finalize_disparity(_np(d), tuple(padded), None if size is None else tuple(size), mode), d))
    monkeypatch.setattr(F_rsm, "warp_by_flow_map", lambda im, fl: _t(oracle.warp_by_flow_map(_np(im), _np(fl)), im))
    def v4_volume_cpu(l, r, conv3d, volume11, d, op_dtype=None):      # the kernels' algebra restated with torch CPU ops
        from test_v4_volume_math import decomposed_volume
        with torch.no_grad():
            return decomposed_volume(F_rsm.pack_v4_weights(conv3d, volume11, torch.float32), l, r, d)
    monkeypatch.setattr(F_rsm, "v4_cost_volume", v4_volume_cpu)
    monkeypatch.setattr(F_rsm, "shift_interweave_volume", lambda l, r, d: _t(oracle.shift_interweave_volume(_np(l), _np(r), d), l))
    sys.path.insert(0, REF)
    yield
    from realtime_stereo_matcher_b200 import unpatch_reference
    unpatch_reference()
    sys.path.remove(REF)


@pytest.mark.parametrize("cfg_name,size,fuse", [
    ("stereo_net_config.json", (48, 96), True), ("stereo_net_config.json", (48, 96), False),
    ("stereo_net_config_v2.json", (48, 96), True), ("stereo_net_config_v3.json", (64, 128), True),
    ("disp_net_c_config.json", (64, 128), True), ("stereo_net_config_v4.json", (32, 224), True),
])
def test_patched_forward_matches_reference(oracle_backed, cfg_name, size, fuse):
    import model as ref_model
    from realtime_stereo_matcher_b200 import patch_reference, unpatch_reference
    cfg = json.load(open(os.path.join(REF, "configure", cfg_name)))
    torch.manual_seed(1234)
    net = ref_model.build_model(cfg["model"]).eval()
    g = torch.Generator().manual_seed(7)
    left = torch.rand((1, 3) + size, generator=g) * 255.0
    right = torch.roll(left, shifts=-3, dims=3)
    with torch.no_grad():
        want = net(left, right)
        done = patch_reference(fuse=fuse)
        assert done["functions"] and (done["forwards"] or not fuse)
        got = net(left, right)
        unpatch_reference()
        again = net(left, right)
    assert len(got) == len(want)
    for a, b, c in zip(got, want, again):
        assert a.shape == b.shape
        torch.testing.assert_close(a, b, atol=2e-3, rtol=1e-4)
        assert torch.equal(b, c)    # unpatch restores the reference bit for bit


def test_patch_lists_everything(oracle_backed):
    from realtime_stereo_matcher_b200 import patch_reference
    done = patch_reference(fuse=True)
    assert len(done["functions"]) == 10 and len(done["classes"]) == 4 and len(done["forwards"]) == 4
    import cost_volume.groupwise as cg
    import realtime_stereo_matcher_b200 as rsm
    assert cg.TorchGroupwiseCost is rsm.TorchGroupwiseCost
    import loss as ref_loss_pkg
    import loss.loss as ref_loss
    assert len(done["loss"]) == 3
    assert ref_loss.SequenceLoss is rsm.SequenceLoss and ref_loss.get_flow_map_metrics is rsm.get_flow_map_metrics
    assert isinstance(ref_loss_pkg.build_loss_function({"type": "SequenceLoss", "parameters": {}}), rsm.SequenceLoss)
    rsm.unpatch_reference()
    assert ref_loss.SequenceLoss is not rsm.SequenceLoss and ref_loss.SequenceLoss.__module__ == "loss.loss"
