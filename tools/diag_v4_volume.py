"""GPU diagnostic / timing of rsm_v4_volume_fwd against the reference loop (cuDNN)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import ref_loader
import realtime_stereo_matcher_b200 as rsm
import test_gpu_v4_volume as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
net = T.net.__wrapped__() if hasattr(T.net, "__wrapped__") else None
if net is None:
    ref = ref_loader.load()
    torch.manual_seed(1234)
    net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).cuda().eval()
for shape in [(1, 4, 64, 8), (2, 16, 64, 48), (1, 7, 203, 48)]:
    B, H, W, D = shape
    g = torch.Generator().manual_seed(1)
    L = torch.randn((B, 32, H, W), generator=g).cuda(); R = torch.randn((B, 32, H, W), generator=g).cuda()
    with torch.no_grad():
        want = T.reference_volume(net, L, R, D)
        for swap in ("0",):
            os.environ["RSM_V4_SWAP_DESC"] = swap
            try:
                got = rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, D)
                torch.cuda.synchronize()
                err = (got - want).abs()
                print(shape, "swap", swap, "max err", float(err.max()), "mean err", float(err.mean()), "scale", float(want.abs().max()),
                      "mean", float(want.abs().mean()), "nan", int(torch.isnan(got).sum()))
                if swap == "0":
                    e = err.amax(dim=(0, 2))      # (D, W)
                    print("   per-d max err", [round(float(v), 4) for v in e.amax(1)[:8]], " worst x per d0", int(e[0].argmax()))
            except Exception as ex:
                print(shape, "swap", swap, "FAILED", repr(ex)[:300])
                raise
os.environ["RSM_V4_SWAP_DESC"] = os.environ.get("V4_BEST", "0")
# timing at the v4 feature shape of 384x1248
for B in (1, 8):
    L = torch.randn((B, 32, 96, 312), device="cuda"); R = torch.randn((B, 32, 96, 312), device="cuda")
    with torch.no_grad():
        for _ in range(2):
            rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, 48)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, 48)
        b.record(); torch.cuda.synchronize()
        print(f"B={B}: fused v4 volume {a.elapsed_time(b) / 5:.3f} ms per call, {a.elapsed_time(b) / 5 / B:.3f} ms per pair")
        if B == 1:
            torch.backends.cudnn.allow_tf32 = True
            T.reference_volume(net, L, R, 48); torch.cuda.synchronize()
            t0 = time.perf_counter(); T.reference_volume(net, L, R, 48); torch.cuda.synchronize()
            print(f"      reference loop (cuDNN, TF32 allowed): {(time.perf_counter() - t0) * 1e3:.2f} ms per pair")
