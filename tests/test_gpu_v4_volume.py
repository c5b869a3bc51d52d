"""GPU suite: rsm_v4_volume_fwd against the reference's own per-disparity loop
(model/mobile_stereo_net_v4.py:443-458) run with the real MobileStereoNetV4 sub-modules (cuDNN, strict fp32) on the device."""
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref missing: run tools/install_ref.py")]


@pytest.fixture(scope="module")
def net():
    ref = ref_loader.load()
    torch.manual_seed(1234)
    net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).cuda().eval()
    g = torch.Generator().manual_seed(5)
    for m in (net.conv3d[1], net.conv3d[4], net.conv3d[7], net.volume11[0][1]):
        m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.1).cuda())
        m.running_var.copy_((1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g)).cuda())
        m.weight.data.copy_((1.0 + 0.1 * torch.randn(m.weight.shape, generator=g)).cuda())
        m.bias.data.copy_((0.05 * torch.randn(m.bias.shape, generator=g)).cuda())
    return net


def reference_volume(net, featL, featR, D):
    v4 = ref_loader.load().v4
    B, C, H, W = featL.shape
    vol = featL.new_zeros([B, D, H, W])
    for i in range(D):
        x = v4.interweave_tensors(featL[:, :, :, i:], featR[:, :, :, : W - i])
        vol[:, i, :, i:] = net.volume11(torch.squeeze(net.conv3d(torch.unsqueeze(x, 1)), 2))[:, 0]
    return vol


# (B, H, W, D): tile tails (W % 128, W % 8 != 0), D > W - 128, a single row, v4's training crop (60 x 80 features)
SHAPES = [(2, 16, 64, 48), (1, 7, 203, 48), (1, 1, 130, 20), (2, 60, 80, 48), (1, 24, 312, 48)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_v4_volume_matches_reference_loop(net, shape, dtype):
    import realtime_stereo_matcher_b200 as rsm
    B, H, W, D = shape
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator().manual_seed(B * 1000 + W)
        featL = torch.randn((B, 32, H, W), generator=g).cuda().to(dtype)
        featR = torch.randn((B, 32, H, W), generator=g).cuda().to(dtype)
        with torch.no_grad():
            want = reference_volume(net, featL.float(), featR.float(), D)       # fp32 truth on the same rounded inputs
            got = rsm.v4_cost_volume(featL, featR, net.conv3d, net.volume11, D)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert got.shape == want.shape and got.dtype == dtype
    assert torch.isfinite(got).all()
    xs = torch.arange(W, device="cuda").view(1, 1, 1, W)
    ds = torch.arange(D, device="cuda").view(1, D, 1, 1)
    assert float(got.masked_select(xs < ds).abs().max() if (xs < ds).any() else 0.0) == 0.0      # x < d is exactly zero
    scale = float(want.abs().max())
    assert scale > 0.05
    err = (got.float() - want).abs()
    # 16-bit operands (fp16: 11-bit significands = TF32; bf16: 8), fp32 accumulation over K = 72 / 576 / 576
    tol_max, tol_mean = (6e-2, 1e-2) if dtype == torch.bfloat16 else (1.5e-2, 2e-3)
    assert float(err.max()) <= tol_max * scale, (float(err.max()), scale)
    assert float(err.mean()) <= tol_mean * float(want.abs().mean()) + 1e-6, (float(err.mean()), float(want.abs().mean()))


def test_v4_volume_error_behaviour(net):
    import realtime_stereo_matcher_b200 as rsm
    l = torch.randn((1, 16, 8, 40), device="cuda")
    with pytest.raises(RuntimeError):                       # C != 32
        rsm.v4_cost_volume(l, l, net.conv3d, net.volume11, 8)
    l = torch.randn((1, 32, 8, 40), device="cuda")
    net.train()
    try:
        with pytest.raises(RuntimeError, match="eval mode"):
            rsm.v4_cost_volume(l, l, net.conv3d, net.volume11, 8)
    finally:
        net.eval()
    assert rsm.v4_cost_volume(l[:0], l[:0], net.conv3d, net.volume11, 8).shape == (0, 8, 8, 40)
