#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference (babiking/realtime_stereo_matcher) has no tests or golden vectors of its
own for the cost-volume / disparity-regression path, so parity is
pinned against its own outputs: this script imports ``cost_volume.*`` and ``model.*``
from /root/reference, runs them on small seeded inputs on the CPU (torch 2.11.0+cu128)
and stores inputs, outputs and autograd gradients as ``*.npz``.  The fixtures travel to
the GPU box; /root/reference does not.  16-bit tensors are stored up-cast to float32
(exact) with their dtype recorded in ``meta``.
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("RSM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from cost_volume.concatenate import TorchConcatenateCost  # noqa: E402
from cost_volume.groupwise import TorchGroupwiseCost  # noqa: E402
from cost_volume.inner_product import TorchInnerProductCost  # noqa: E402
from cost_volume.interweave import TorchInterweaveCost  # noqa: E402
import model as ref_model  # noqa: E402
import model.mobile_disp_net_c as ref_dispc  # noqa: E402
import model.mobile_stereo_net as ref_v1  # noqa: E402
import model.mobile_stereo_net_v2 as ref_v2  # noqa: E402
import model.mobile_stereo_net_v3 as ref_v3  # noqa: E402
import model.mobile_stereo_net_v4 as ref_v4  # noqa: E402

DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def npy(t):
    t = t.detach()
    if t.dtype in (torch.float16, torch.bfloat16):
        t = t.float()
    return t.contiguous().numpy()


def gen(shape, kind, dtype, g):
    if kind == "dyadic":  # k/8, k in [-8, 8]: fp32 sums are order independent
        t = torch.randint(-8, 9, shape, generator=g).float() / 8.0
    else:
        t = torch.randn(shape, generator=g)
    return t.to(dtype)


def save(name, meta, **arrays):
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(f"{name}: {len(arrays) - 1} arrays")


def pair_with_grads(fn, l, r, g, grad=True):
    """Run fn(l, r) with autograd and a seeded upstream gradient."""
    l = l.clone().requires_grad_(grad)
    r = r.clone().requires_grad_(grad)
    out = fn(l, r)
    res = {"out": npy(out)}
    if grad:
        gout = torch.randn(out.shape, generator=g).to(out.dtype)
        out.backward(gout)
        res.update(gout=npy(gout), gleft=npy(l.grad), gright=npy(r.grad))
    return res


# (tag, N, C, H, W, D, G) -- edge cases from SURVEY.md 8c: D > W, D = 1, W = 1, C = G, odd D
SHAPES = [
    ("base", 2, 8, 5, 13, 6, 4),
    ("oddD", 1, 6, 3, 17, 7, 3),
    ("DgtW", 1, 4, 2, 5, 9, 2),
    ("D1", 2, 4, 3, 8, 1, 4),
    ("W1", 1, 4, 3, 1, 3, 1),
    ("CeqG", 1, 8, 4, 12, 4, 8),
    ("vec", 1, 16, 4, 32, 8, 4),
]


def volume_goldens():
    for kind in ("normal", "dyadic"):
        for tag, n, c, h, w, d, ng in SHAPES:
            for dname in ("fp32", "fp16", "bf16"):
                if dname != "fp32" and tag not in ("base", "vec"):
                    continue
                dtype = DT[dname]
                g = torch.Generator().manual_seed(1234)
                l = gen((n, c, h, w), kind, dtype, g)
                r = gen((n, c, h, w), kind, dtype, g)
                meta = dict(tag=tag, kind=kind, dtype=dname, N=n, C=c, H=h, W=w, D=d, G=ng)
                base = {"left": npy(l), "right": npy(r)}
                out = {}
                for op, fn in (
                    ("concat", TorchConcatenateCost(d)),
                    ("interweave", TorchInterweaveCost()),
                    ("interweave_v4", ref_v4.interweave_tensors),
                    ("inner", TorchInnerProductCost(d)),
                    ("corr_mean", lambda a, b: ref_dispc.make_correlation_volume(a, b, d)),
                    ("groupwise", TorchGroupwiseCost(ng, d)),
                    ("difference", lambda a, b: ref_v1.make_cost_volume(a, b, d)),
                ):
                    res = pair_with_grads(fn, l, r, g)
                    for k, v in res.items():
                        out[f"{op}.{k}"] = v
                save(f"vol_{tag}_{kind}_{dname}", meta, **base, **out)


def noncontig_golden():
    """v4 passes width-cropped, non-contiguous slices (mobile_stereo_net_v4.py:446)."""
    g = torch.Generator().manual_seed(99)
    fl = gen((2, 6, 4, 20), "normal", torch.float32, g)
    fr = gen((2, 6, 4, 20), "normal", torch.float32, g)
    i = 3
    a, b = fl[:, :, :, i:], fr[:, :, :, :-i]
    save("noncontig_interweave", dict(i=i), featL=npy(fl), featR=npy(fr),
         out=npy(ref_v4.interweave_tensors(a, b)),
         inner=npy(TorchInnerProductCost(5)(a, b)),
         concat=npy(TorchConcatenateCost(5)(a, b)))


def regression_goldens():
    for tag, n, d, h, w, scale in (("base", 2, 12, 5, 9, 4.0), ("D1", 1, 1, 3, 4, 1.0),
                                   ("wide", 1, 48, 4, 33, 8.0), ("flat", 1, 7, 2, 5, 0.0)):
        for dname in ("fp32", "fp16", "bf16"):
            g = torch.Generator().manual_seed(4321)
            cost = (torch.randn((n, d, h, w), generator=g) * scale).to(DT[dname])
            c1 = cost.clone().requires_grad_(True)
            # DispNetC form: logits in, (N,1,H,W) out -- mobile_disp_net_c.py:208-220
            e1 = ref_dispc.disparity_regression(c1, d)
            gout = torch.randn(e1.shape, generator=g).to(e1.dtype)
            e1.backward(gout)
            # v4 form: probabilities in, (N,H,W) out -- mobile_stereo_net_v4.py:10-14
            e2 = ref_v4.disparity_regression(F.softmax(cost, dim=1), d)
            save(f"regress_{tag}_{dname}", dict(tag=tag, dtype=dname, N=n, D=d, H=h, W=w),
                 cost=npy(cost), e_keepdim=npy(e1), e=npy(e2), gout=npy(gout), gcost=npy(c1.grad),
                 argmin=torch.argmin(cost.float(), dim=1).numpy(),
                 argmax=torch.argmax(cost.float(), dim=1).numpy())
    # ties and NaN: first occurrence wins, NaN counts as the extremum
    cost = torch.tensor([[3.0, 1.0, 1.0, 2.0], [float("nan"), 0.0, -1.0, float("nan")],
                         [5.0, 5.0, 5.0, 5.0], [1.0, float("nan"), float("nan"), 0.0]])
    cost = cost.t().reshape(1, 4, 1, 4).contiguous()
    save("regress_ties_nan", dict(N=1, D=4, H=1, W=4), cost=npy(cost),
         argmin=torch.argmin(cost, dim=1).numpy(), argmax=torch.argmax(cost, dim=1).numpy())


def tail_goldens():
    """v4 head: interpolate(trilinear) -> softmax -> disparity_regression,
    mobile_stereo_net_v4.py:511-518."""
    for tag, b, dc, hc, wc, od, oh, ow in (("x4", 2, 6, 3, 5, 24, 12, 20),
                                           ("v4like", 1, 48, 4, 6, 192, 16, 24),
                                           ("uneven", 1, 5, 3, 4, 13, 7, 10)):
        g = torch.Generator().manual_seed(77)
        cost = (torch.randn((b, dc, hc, wc), generator=g) * 3.0).requires_grad_(True)
        fine = F.interpolate(cost.unsqueeze(1), [od, oh, ow], mode="trilinear").squeeze(1)
        pred = ref_v4.disparity_regression(F.softmax(fine, dim=1), od)
        gout = torch.randn(pred.shape, generator=g)
        pred.backward(gout)
        save(f"tail_{tag}", dict(tag=tag, B=b, Dc=dc, Hc=hc, Wc=wc, D=od, H=oh, W=ow),
             cost=npy(cost), fine=npy(fine), pred=npy(pred), gout=npy(gout), gcost=npy(cost.grad),
             argmin=torch.argmin(fine.detach(), dim=1).numpy(),
             argmax=torch.argmax(fine.detach(), dim=1).numpy())


def record(module, name, store, limit=None):
    """Wrap module-global function ``name`` so calls are recorded (the models look these
    functions up as module globals at call time -- SURVEY.md 8b)."""
    orig = getattr(module, name)

    def wrapper(*a, **k):
        out = orig(*a, **k)
        if limit is None or len(store) < limit:
            args = list(a) + list(k.values())
            store.append(([x.detach().clone() if torch.is_tensor(x) else x for x in args], out.detach().clone()))
        return out

    setattr(module, name, wrapper)
    return orig


def model_callsite_goldens():
    cfgdir = os.path.join(REF, "configure")
    torch.manual_seed(1234)
    g = torch.Generator().manual_seed(1234)
    limg = torch.rand((1, 3, 64, 128), generator=g) * 255.0
    rimg = torch.roll(limg, shifts=-5, dims=3)

    # --- v1: make_cost_volume call site mobile_stereo_net.py:140, regression :144-147
    cfg = json.load(open(os.path.join(cfgdir, "stereo_net_config.json")))
    net = ref_model.build_model(cfg["model"]).eval()
    calls, filt, reg = [], [], []
    orig = record(ref_v1, "make_cost_volume", calls)
    h1 = net.cost_filter.register_forward_hook(lambda m, i, o: filt.append(o.detach().clone()))
    h2 = net.refine_layer[0].register_forward_pre_hook(lambda m, i: reg.append(i[0].detach().clone()))
    with torch.no_grad():
        outs = net(limg, rimg)
    h1.remove(), h2.remove()
    setattr(ref_v1, "make_cost_volume", orig)
    (lf, rf, md), vol = calls[0]
    save("callsite_v1", dict(max_disp=int(md)), lf=npy(lf), rf=npy(rf), volume=npy(vol),
         filtered=npy(filt[0].squeeze(1)), regressed=npy(reg[0]), final=npy(outs[-1]))

    # --- DispNetC: make_correlation_volume call site mobile_disp_net_c.py:365-367
    cfg = json.load(open(os.path.join(cfgdir, "disp_net_c_config.json")))
    net = ref_model.build_model(cfg["model"]).eval()
    calls = []
    orig = record(ref_dispc, "make_correlation_volume", calls)
    with torch.no_grad():
        net(limg, rimg)
    setattr(ref_dispc, "make_correlation_volume", orig)
    (lf, rf, md), vol = calls[0]
    save("callsite_dispnetc", dict(max_disp=int(md)), lf=npy(lf), rf=npy(rf), volume=npy(vol))

    # --- v4: interweave_tensors call sites :446/:453, tail :511-518
    cfg = json.load(open(os.path.join(cfgdir, "stereo_net_config_v4.json")))
    net = ref_model.build_model(cfg["model"]).eval()
    calls, cls3 = [], []
    orig = record(ref_v4, "interweave_tensors", calls, limit=4)
    h = net.classif3.register_forward_hook(lambda m, i, o: cls3.append(o.detach().clone()))
    # v4 needs W/4 - 47 >= 3 (Conv3d kernel width) -> 64 x 256 image
    limg4 = torch.rand((1, 3, 64, 256), generator=g) * 255.0
    rimg4 = torch.roll(limg4, shifts=-7, dims=3)
    with torch.no_grad():
        outs = net(limg4, rimg4)
    h.remove()
    setattr(ref_v4, "interweave_tensors", orig)
    arrays = {}
    for k, ((a, b), o) in enumerate(calls):
        arrays[f"iw{k}.a"], arrays[f"iw{k}.b"], arrays[f"iw{k}.out"] = npy(a), npy(b), npy(o)
    save("callsite_v4", dict(maxdisp=int(net.maxdisp), n_iw=len(calls), H=64, W=256),
         cost3=npy(cls3[0]), final=npy(outs[-1]), **arrays)


def warp_goldens():
    """warp_by_flow_map (refinement warp): v2 / v3 / tools copies are the same function; the
    flow fields push samples out of the image on both sides; 1- and 2-channel flows."""
    g = torch.Generator().manual_seed(4321)
    for tag, n, c, h, w, cf in (("base", 2, 3, 6, 20, 1), ("two", 1, 4, 5, 9, 2), ("feat", 1, 8, 12, 40, 1)):
        image = torch.randn((n, c, h, w), generator=g).requires_grad_(True)
        flow = (torch.randn((n, cf, h, w), generator=g) * 3.0 + 2.0).requires_grad_(True)
        out = ref_v2.warp_by_flow_map(image, flow)
        assert torch.equal(out, ref_v3.warp_by_flow_map(image, flow))
        gout = torch.randn(out.shape, generator=g)
        out.backward(gout)
        save(f"warp_{tag}", dict(tag=tag, N=n, C=c, H=h, W=w, CF=cf), image=npy(image), flow=npy(flow), out=npy(out),
             gout=npy(gout), gimage=npy(image.grad), gflow=npy(flow.grad))


def prepost_goldens():
    """Pre / post steps exactly as the reference models execute them.  v1: the normalised + padded
    images are what `feature_extractor` receives (mobile_stereo_net.py:121-130), the post step maps each RefineNet
    output to the returned full-resolution map (:154-159); gradients by autograd through the model's own graph.
    DispNetC: `disparity_interpolate` (mobile_disp_net_c.py:223-234) followed by the crop + negate of :408-411."""
    cfgdir = os.path.join(REF, "configure")
    torch.manual_seed(99)
    g = torch.Generator().manual_seed(99)
    cfg = json.load(open(os.path.join(cfgdir, "stereo_net_config.json")))
    net = ref_model.build_model(cfg["model"]).eval()
    limg = (torch.rand((1, 3, 52, 90), generator=g) * 255.0).requires_grad_(True)
    rimg = torch.roll(limg.detach(), shifts=-3, dims=3)
    prepared, refined = [], []
    hooks = [net.feature_extractor.register_forward_pre_hook(lambda m, i: prepared.append(i[0]))]
    hooks += [r.register_forward_hook(lambda m, i, o: refined.append(o)) for r in net.refine_layer]
    outs = net(limg, rimg)
    for hk in hooks:
        hk.remove()
    gprep = torch.randn(prepared[0].shape, generator=g)
    (glimg,) = torch.autograd.grad(prepared[0], limg, gprep, retain_graph=True)
    arrays = dict(limg=npy(limg), rimg=npy(rimg), prep_l=npy(prepared[0]), prep_r=npy(prepared[1]), gprep=npy(gprep),
                  glimg=npy(glimg))
    for k, (x, o) in enumerate(zip(refined, outs)):
        go = torch.randn(o.shape, generator=g)
        (gx,) = torch.autograd.grad(o, x, go, retain_graph=True)
        arrays[f"refined{k}"], arrays[f"final{k}"], arrays[f"gfinal{k}"], arrays[f"grefined{k}"] = npy(x), npy(o), npy(go), npy(gx)
    save("prepost_v1", dict(align=int(net.align), H=52, W=90, n_out=len(outs), mode="nearest"), **arrays)

    # DispNetC: six scales of one 52 x 90 frame padded to 64 x 128 (align 2**6), the last already at full size
    hp, wp, h, w = 64, 128, 52, 90
    arrays = {}
    shapes = [(1, 2), (2, 4), (4, 8), (8, 16), (32, 64), (64, 128)]
    for k, (hs, ws) in enumerate(shapes):
        disp = (torch.randn((2, 1, hs, ws), generator=g) * 4.0).requires_grad_(True)
        out = -1.0 * ref_dispc.disparity_interpolate(disp, (hp, wp))[:, :, :h, :w]
        go = torch.randn(out.shape, generator=g)
        out.backward(go)
        arrays[f"disp{k}"], arrays[f"out{k}"], arrays[f"gout{k}"], arrays[f"gdisp{k}"] = npy(disp), npy(out), npy(go), npy(disp.grad)
    save("prepost_dispnetc", dict(Hp=hp, Wp=wp, H=h, W=w, n=len(shapes), mode="bilinear"), **arrays)


def loss_goldens():
    """SequenceLoss / get_flow_map_metrics executed from the reference's loss/loss.py: predictions at
    1/4, 1/2 and full size plus a fractional ratio, a ground truth that exceeds max_flow_magnitude in places and a
    random validity mask; gradients of the loss with respect to every prediction by autograd."""
    from loss.loss import SequenceLoss as RefSequenceLoss, get_flow_map_metrics as ref_metrics
    g = torch.Generator().manual_seed(2024)
    for tag, (n, h, w), sizes, max_flow in (("pyramid", (2, 24, 40), [(6, 10), (12, 20), (24, 40)], 30.0),
                                             ("fractional", (1, 20, 30), [(7, 11), (20, 30)], 700.0),
                                             ("single", (3, 9, 13), [(9, 13)], 5.0)):
        gt = torch.randn((n, 1, h, w), generator=g) * 20.0
        valid = (torch.rand((n, h, w), generator=g) > 0.3).float()
        preds = [(F.interpolate(gt, s) * (s[1] / w) + torch.randn((n, 1) + s, generator=g) * 1.5).requires_grad_(True)
                 for s in sizes]
        loss = RefSequenceLoss(loss_gamma=0.9, max_flow_magnitude=max_flow)(preds, gt, valid)
        loss.backward()
        metrics = ref_metrics(gt, preds[-1].detach(), valid)
        arrays = dict(gt=npy(gt), valid=npy(valid), loss=npy(loss))
        for k, p in enumerate(preds):
            arrays[f"pred{k}"], arrays[f"gpred{k}"] = npy(p), npy(p.grad)
        save(f"loss_{tag}", dict(tag=tag, n_preds=len(preds), gamma=0.9, max_flow=max_flow, metrics=metrics), **arrays)


def pfm_goldens():
    """PFM files written by the reference's tools/pfm_file_io.py (the bytes themselves) and what its reader returns."""
    import tempfile
    from tools.pfm_file_io import read_pfm_file as ref_read, write_pfm_file as ref_write
    rng = np.random.default_rng(7)
    arrays, meta = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        for k, (shape, scale) in enumerate((((5, 7), 1), ((4, 6, 1), 2.5), ((3, 5, 3), 1.0), ((1, 1), 0.125))):
            img = (rng.standard_normal(shape) * 50).astype(np.float32)
            path = os.path.join(tmp, f"g{k}.pfm")
            ref_write(path, img, scale)
            data, rscale = ref_read(path)
            arrays[f"img{k}"], arrays[f"bytes{k}"] = img, np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
            arrays[f"read{k}"] = np.ascontiguousarray(data)
            meta.append(dict(scale=scale, read_scale=rscale))
        # the call test_stereo.py:133 makes
        disp = (rng.random((6, 9)) * 192).astype(np.float32)
        path = os.path.join(tmp, "disp.pfm")
        ref_write(path, np.flipud(disp), 1.0)
        arrays["disp"], arrays["disp_bytes"] = disp, np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    save("pfm_files", dict(cases=meta), **arrays)


if __name__ == "__main__":
    torch.set_num_threads(4)
    if "--only-pfm" in sys.argv:
        pfm_goldens()
        sys.exit(0)
    if "--only-loss" in sys.argv:
        loss_goldens()
        sys.exit(0)
    if "--only-prepost" in sys.argv:
        prepost_goldens()
        sys.exit(0)
    if "--only-warp" in sys.argv:
        warp_goldens()
        sys.exit(0)
    volume_goldens()
    noncontig_golden()
    regression_goldens()
    tail_goldens()
    model_callsite_goldens()
    warp_goldens()
    prepost_goldens()
    loss_goldens()
    pfm_goldens()
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"total fixture bytes: {total}")
