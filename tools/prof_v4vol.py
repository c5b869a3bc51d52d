"""Per-role cycle counters of the two v4 implicit-GEMM kernels (rsm_v4_volume_fwd_profile)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader
import realtime_stereo_matcher_b200 as rsm
from realtime_stereo_matcher_b200 import _lib as L
from realtime_stereo_matcher_b200.functional import pack_v4_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ref = ref_loader.load()
torch.manual_seed(1234)
net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).cuda().eval()
l = torch.randn((B, 32, 96, 312), device="cuda"); r = torch.randn((B, 32, 96, 312), device="cuda")
lib = L.load()
pk = pack_v4_weights(net.conv3d, net.volume11, torch.float16)
out = torch.empty((B, 48, 96, 312), device="cuda")
work = torch.empty(lib.rsm_v4_volume_workspace(B, 96, 312, 48), dtype=torch.uint8, device="cuda")
wts = L.RsmV4Weights(*(pk[k].data_ptr() for k in ("w1", "t1", "w2", "t2", "w3", "t3", "w11", "t11")))
for rep in range(3):
    prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.rsm_v4_volume_fwd_profile(L.feat(l), L.feat(r), wts, out.data_ptr(), work.data_ptr(), B, 32, 96, 312, 48, 0, 1, 0,
                                          L.stream_ptr(0), prof.data_ptr()), "profile")
    b.record(); torch.cuda.synchronize()
p = prof.cpu().tolist()
ms = a.elapsed_time(b)
for name, o, ncta in (("layer 2 (GEN)", 0, min(148, B * 48 * 2 * 3)), ("layer 3 (TMA)", 16, min(148, B * 48 * 3))):
    rows = p[o + 11] / ncta if p[o + 11] else float("nan")
    f = lambda v: f"{v / ncta / max(rows, 1):8.0f}"
    print(f"{name}: rows/CTA {rows:.0f}; cycles per row (CTA average): issuer total {f(p[o+2])} = wait rows {f(p[o+0])} + wait acc {f(p[o+1])} + issue {f(p[o+2]-p[o+0]-p[o+1])}")
    print(f"     producer g0: total {f(p[o+5])} wait slot {f(p[o+3])} work {f(p[o+4])}")
    print(f"     producer g1: wait slot {f(p[o+6])} work {f(p[o+7])} | epilogue: total {f(p[o+10])} wait acc {f(p[o+9])}")
print(f"whole call {ms:.3f} ms for {B} pairs")
