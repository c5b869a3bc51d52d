"""CPU suite (build container: needs the reference): the algebra behind rsm_v4_volume_fwd -- layer 1 split into
left-only / right-only maps computed once per pair plus two edge maps that restore the cropped tensor's zero padding,
BatchNorm folded, layers 2 / 3 as 3x3 convolutions over depth blocks with the weights PACKED as the kernels read
them, x < d re-zeroed after every layer -- restated with torch CPU ops and checked against the reference's own
per-disparity loop (model/mobile_stereo_net_v4.py:443-458) on the real module.  The CUDA kernels are checked against
the same loop on the GPU (tests/test_gpu_v4_volume.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref missing")


def reference_volume(net, featL, featR):
    """The loop of MobileStereoNetV4.forward, verbatim semantics (:443-458)."""
    B, C, H, W = featL.shape
    vol = featL.new_zeros([B, 1, net.volume_size, H, W])
    v4 = ref_loader.load().v4
    for i in range(net.volume_size):
        if i > 0:
            x = v4.interweave_tensors(featL[:, :, :, i:], featR[:, :, :, :-i])
        else:
            x = v4.interweave_tensors(featL, featR)
        x = net.volume11(torch.squeeze(net.conv3d(torch.unsqueeze(x, 1)), 2))
        vol[:, :, i, :, i:] = x
    return vol.squeeze(1)


def decomposed_volume(pk, featL, featR, D):
    """What K1 / K2 / K3 of csrc/rsm_v4vol.cu compute, in fp32 torch ops, from the PACKED weights."""
    B, C, H, W = featL.shape
    w1 = pk["w1"]                                            # (16,8,3,3) [o][kd][dy][dx]
    wL = w1[:, 0::2].repeat(8, 1, 1, 1)                      # (128,4,3,3): group k -> left channels 4k..4k+3
    wR = w1[:, 1::2].repeat(8, 1, 1, 1)
    PL = F.conv2d(featL, wL, padding=1, groups=8) + pk["t1"].repeat(8).view(1, -1, 1, 1)
    PR = F.conv2d(featR, wR, padding=1, groups=8)
    wLe, wRe = wL.clone(), wR.clone()
    wLe[..., 1:] = 0                                         # only the dx = 0 column (reads x - 1)
    wRe[..., :2] = 0                                         # only the dx = 2 column (reads x' + 1)
    EL = F.conv2d(featL, wLe, padding=1, groups=8)
    ER = F.conv2d(featR, wRe, padding=1, groups=8)

    def unpack(p, cin_total):                                # (9,8,co,8) -> (co, 64, 3, 3) with kk = chunk*8 + e
        co = p.shape[2]
        return p.float().permute(2, 1, 3, 0).reshape(co, 64, 3, 3)

    W2, W3 = unpack(pk["w2"], 64), unpack(pk["w3"], 64)
    out = featL.new_zeros((B, D, H, W))
    xs = torch.arange(W)
    for d in range(D):
        keep = (xs >= d).float().view(1, 1, 1, W)
        PRs = torch.zeros_like(PR)
        ERs = torch.zeros_like(ER)
        PRs[..., d:] = PR[..., : W - d]
        ERs[..., d:] = ER[..., : W - d]
        a1 = PL + PRs
        a1[..., d] -= EL[..., d]
        a1[..., W - 1] -= ERs[..., W - 1]
        a1 = F.relu(a1) * keep                              # channel = k*16 + o, k = 4j + kd  ->  block j = [64j, 64j+64)
        a2 = torch.cat([F.relu(F.conv2d(a1[:, 64 * j:64 * j + 64], W2, padding=1) + pk["t2"].view(1, -1, 1, 1)) for j in range(2)], 1) * keep
        a3 = F.relu(F.conv2d(a2, W3, padding=1) + pk["t3"].view(1, -1, 1, 1))
        v = F.relu((a3 * pk["w11"].view(1, -1, 1, 1)).sum(1) + pk["t11"]) * keep[:, 0]
        out[:, d] = v
    return out


def test_decomposition_matches_reference_loop():
    from realtime_stereo_matcher_b200.functional import pack_v4_weights
    ref = ref_loader.load()
    torch.manual_seed(1234)
    net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).eval()
    g = torch.Generator().manual_seed(5)
    for m in (net.conv3d[1], net.conv3d[4], net.conv3d[7], net.volume11[0][1]):
        m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
        m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
        m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
        m.bias.data.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    net.volume_size = 9                                       # a short loop is enough for the algebra
    featL = torch.randn((2, 32, 5, 21), generator=g)
    featR = torch.randn((2, 32, 5, 21), generator=g)
    pk = pack_v4_weights(net.conv3d, net.volume11, torch.float32)
    with torch.no_grad():
        want = reference_volume(net, featL, featR)
        got = decomposed_volume(pk, featL, featR, net.volume_size)
    assert float(want.abs().max()) > 0.1
    torch.testing.assert_close(got, want, atol=2e-5, rtol=1e-4)
