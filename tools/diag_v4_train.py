"""GPU diagnostic: which gradients of the patched MobileStereoNetV4 training step differ from the unpatched run."""
import contextlib, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader
import realtime_stereo_matcher_b200 as rsm
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_models as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ref = ref_loader.load()

# 1) v4 head backward alone
torch.manual_seed(0)
c = (torch.randn((2, 48, 16, 64), device="cuda") * 3)
g = torch.randn((2, 64, 256), device="cuda")
c1 = c.clone().requires_grad_(True)
p1 = ref.v4.disparity_regression(F.softmax(F.interpolate(c1.unsqueeze(1), [192, 64, 256], mode="trilinear").squeeze(1), 1), 192)
p1.backward(g)
c2 = c.clone().requires_grad_(True)
p2 = rsm.v4_head(c2, 192, 64, 256)
p2.backward(g)
print("v4 head fwd err", float((p1 - p2).abs().max()), "bwd rel err", float((c1.grad - c2.grad).abs().max() / c1.grad.abs().max()))

# 2) interweave backward on cropped views
l = torch.randn((2, 32, 16, 64), device="cuda", requires_grad=True)
r = torch.randn((2, 32, 16, 64), device="cuda", requires_grad=True)
go = torch.randn((2, 64, 16, 64 - 5), device="cuda")
ref.v4.interweave_tensors(l[:, :, :, 5:], r[:, :, :, :-5]).backward(go)
gl, gr = l.grad.clone(), r.grad.clone(); l.grad = None; r.grad = None
rsm.interweave_tensors(l[:, :, :, 5:], r[:, :, :, :-5]).backward(go)
print("interweave bwd err", float((gl - l.grad).abs().max()), float((gr - r.grad).abs().max()))

# 3) per-parameter gradient error of the training step, patched piece by piece
cfg_name, size, batch = T.MODELS[3]
left, right = T.stereo_pair(size, batch)
gen = torch.Generator().manual_seed(11)
flow_gt = -(5.0 + torch.rand((batch, 1) + size, generator=gen)).cuda()
valid = (torch.rand((batch,) + size, generator=gen) > 0.2).float().cuda()

def step(mode):
    net, cfg = T.build(ref, cfg_name, train=True)
    if mode == "patched":
        ctx = T.patched()
    elif mode == "nofuse":
        ctx = T.patched(fuse=False)
    else:
        ctx = contextlib.nullcontext()
    with ctx:
        loss_fn = ref.loss.build_loss_function(cfg["train"]["loss"])
        preds = net(left, right)
        loss = loss_fn(preds, flow_gt, valid)
        loss.backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}, [p.detach() for p in preds]

l0, g0, p0 = step("ref")
l0b, g0b, p0b = step("ref")
for mode in ("nofuse", "patched"):
    l1, g1, p1 = step(mode)
    num = sum(float((g1[k] - g0[k]).double().pow(2).sum()) for k in g0); den = sum(float(g0[k].double().pow(2).sum()) for k in g0)
    print(mode, "loss", l0, l1, "grad rel", (num / den) ** 0.5, "pred err", [float((a - b).abs().max()) for a, b in zip(p1, p0)])
    worst = sorted(((float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-12)), k) for k in g0), reverse=True)[:8]
    print("  worst:", worst)
num = sum(float((g0b[k] - g0[k]).double().pow(2).sum()) for k in g0); den = sum(float(g0[k].double().pow(2).sum()) for k in g0)
print("ref vs ref (run-to-run)", (num / den) ** 0.5)

# 4) the fused pieces on the tensors of a real (unpatched) training forward
net, cfg = T.build(ref, cfg_name, train=True)
costs = {}
hooks = [getattr(net, f"classif{i}").register_forward_hook(lambda m, i_, o, k=i: costs.__setitem__(k, o.detach())) for i in range(4)]
with torch.no_grad():
    preds = net(left, right)
[h.remove() for h in hooks]
img_ref = (2.0 * (left / 255.0) - 1.0).contiguous()
print("prepare_input bit-exact:", bool(torch.equal(rsm.prepare_input(left), img_ref)))
for k in range(4):
    c = costs[k]
    c1 = c.clone().requires_grad_(True)
    p1 = ref.v4.disparity_regression(F.softmax(F.interpolate(c1.unsqueeze(1), [192, size[0], size[1]], mode="trilinear").squeeze(1), 1), 192)
    gg = torch.randn_like(p1)
    p1.backward(gg)
    c2 = c.clone().requires_grad_(True)
    p2 = rsm.v4_head(c2, 192, size[0], size[1])
    p2.backward(gg)
    print(f"head {k}: cost range [{float(c.min()):.2f}, {float(c.max()):.2f}] fwd err {float((p1 - p2).abs().max()):.5f} vs model pred err {float((-p1.detach().unsqueeze(1) - preds[k]).abs().max()):.5f}"
          f" bwd rel err {float((c1.grad - c2.grad).norm() / c1.grad.norm()):.2e}")

# 5) how sensitive is the UNPATCHED training step to a one-ulp change of its input normalisation?
def step_ref_inputs(l_in, r_in):
    net, cfg = T.build(ref, cfg_name, train=True)
    loss_fn = ref.loss.build_loss_function(cfg["train"]["loss"])
    preds = net(l_in, r_in)
    loss = loss_fn(preds, flow_gt, valid)
    loss.backward()
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}, [p.detach() for p in preds]
ga, pa = step_ref_inputs(left, right)
gb, pb = step_ref_inputs(left * (1 + 2.0 ** -23), right * (1 + 2.0 ** -23))
num = sum(float((ga[k] - gb[k]).double().pow(2).sum()) for k in ga); den = sum(float(ga[k].double().pow(2).sum()) for k in ga)
print("unpatched, inputs perturbed by 1 ulp: grad rel", (num / den) ** 0.5, "pred err", [float((a - b).abs().max()) for a, b in zip(pa, pb)])
# 6) patched (fused) but with the normalisation done by torch on the device (mul by reciprocal, ATen's CUDA div)
import realtime_stereo_matcher_b200.model_functions as mf
orig = mf.prepare_input
mf.prepare_input = lambda img, align=1: (2.0 * (img / 255.0) - 1.0).contiguous()
try:
    l1, g1, p1 = step("patched")
finally:
    mf.prepare_input = orig
num = sum(float((g1[k] - g0[k]).double().pow(2).sum()) for k in g0); den = sum(float(g0[k].double().pow(2).sum()) for k in g0)
print("patched with torch-CUDA normalisation: grad rel", (num / den) ** 0.5, "pred err", [float((a - b).abs().max()) for a, b in zip(p1, p0)])
