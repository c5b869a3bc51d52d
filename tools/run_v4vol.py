"""ncu driver: a few calls of rsm_v4_volume_fwd at the v4 feature shape of 384x1248 (B pairs), random weights."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader
import realtime_stereo_matcher_b200 as rsm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ref = ref_loader.load()
torch.manual_seed(1234)
net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).cuda().eval()
L = torch.randn((B, 32, 96, 312), device="cuda"); R = torch.randn((B, 32, 96, 312), device="cuda")
with torch.no_grad():
    for _ in range(reps):
        v = rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, 48)
torch.cuda.synchronize()
print("ok", tuple(v.shape), float(v.abs().mean()))
