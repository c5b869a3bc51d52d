"""GPU suite: the shifted interweave stack (first step) and the batched form of
MobileStereoNetV4's per-disparity volume loop built on it."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import oracle
from golden_io import round_to

pytestmark = pytest.mark.gpu
DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()
    return m


@pytest.mark.parametrize("shape", [(2, 8, 5, 64, 12), (1, 32, 4, 312, 48), (1, 5, 3, 67, 19), (2, 3, 2, 9, 13)])
@pytest.mark.parametrize("dn", ["fp32", "bf16", "fp16"])
def test_shift_interweave_vs_oracle(rsm, shape, dn):
    n, c, h, w, d = shape
    rng = np.random.default_rng(17)
    l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
    L = torch.from_numpy(l).cuda().to(DT[dn]).requires_grad_(dn == "fp32")
    R = torch.from_numpy(r).cuda().to(DT[dn]).requires_grad_(dn == "fp32")
    out = rsm.shift_interweave_volume(L, R, d)
    assert out.shape == (d, n, 2 * c, h, w) and out.is_contiguous()
    np.testing.assert_array_equal(out.detach().float().cpu().numpy(), oracle.shift_interweave_volume(l, r, d))  # bit exact
    if dn == "fp32":
        g = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(torch.from_numpy(g).cuda())
        gl, gr = oracle.shift_interweave_volume_bwd(g)
        np.testing.assert_allclose(L.grad.cpu().numpy(), gl, atol=1e-4)
        np.testing.assert_allclose(R.grad.cpu().numpy(), gr, atol=1e-4)


def test_shift_interweave_equals_reference_loop_inputs(rsm):
    """out[i][..., i:] is exactly what iteration i of the reference loop feeds to its convolutions
    (interweave of the cropped views), checked against our own interweave on the cropped slices."""
    g = torch.Generator(device="cuda").manual_seed(3)
    fl = torch.randn((2, 16, 6, 80), device="cuda", generator=g)
    fr = torch.randn((2, 16, 6, 80), device="cuda", generator=g)
    stack = rsm.shift_interweave_volume(fl, fr, 20)
    for i in (0, 1, 7, 19):
        ref = rsm.interweave_tensors(fl[:, :, :, i:], fr[:, :, :, : 80 - i])
        assert torch.equal(stack[i][..., i:], ref)
        assert stack[i][..., :i].abs().sum().item() == 0.0


class _V4Like(nn.Module):
    """Stand-in with the attribute names and layer shapes of MobileStereoNetV4's volume builder
    (model/mobile_stereo_net_v4.py:317-335); the reference itself is not present on the GPU box."""

    def __init__(self, volume_size=12):
        super().__init__()
        self.volume_size, self.num_groups = volume_size, 1
        self.conv3d = nn.Sequential(
            nn.Conv3d(1, 16, kernel_size=(8, 3, 3), stride=[8, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(16), nn.ReLU(),
            nn.Conv3d(16, 32, kernel_size=(4, 3, 3), stride=[4, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(32), nn.ReLU(),
            nn.Conv3d(32, 16, kernel_size=(2, 3, 3), stride=[2, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(16), nn.ReLU())
        self.volume11 = nn.Sequential(nn.Conv2d(16, 1, 1, 1, 0, bias=False), nn.BatchNorm2d(1), nn.ReLU(inplace=True))


def test_v4_batched_volume_matches_loop(rsm):
    from realtime_stereo_matcher_b200.patch import v4_volume_batched
    torch.manual_seed(0)
    net = _V4Like().cuda().eval()
    for m in net.modules():                      # non-trivial BN statistics
        if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm2d)):
            m.running_mean.normal_(0, 0.2), m.running_var.uniform_(0.5, 1.5), m.weight.data.uniform_(0.5, 1.5), m.bias.data.normal_(0, 0.2)
    g = torch.Generator(device="cuda").manual_seed(9)
    fl = torch.randn((2, 32, 10, 40), device="cuda", generator=g)
    fr = torch.randn((2, 32, 10, 40), device="cuda", generator=g)
    B, C, H, W = fl.shape
    with torch.no_grad():
        want = fl.new_zeros([B, net.volume_size, H, W])
        for i in range(net.volume_size):         # the reference's per-disparity loop (:444-458)
            x = rsm.interweave_tensors(fl[:, :, :, i:], fr[:, :, :, : W - i]).unsqueeze(1)
            want[:, i, :, i:] = net.volume11(net.conv3d(x).squeeze(2)).squeeze(1)
        got = v4_volume_batched(net, fl, fr)
        got_chunked = v4_volume_batched(net, fl, fr, chunk=5)
    torch.testing.assert_close(got, want, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(got_chunked, want, atol=1e-4, rtol=1e-4)
