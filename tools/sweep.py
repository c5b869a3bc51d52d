#!/usr/bin/env python
"""Per-kernel roofline sweep over the BASELINE.json configs.

Each op is timed alone with CUDA events on the launching stream, L2 flushed (a 512 MB memset)
before every timed launch, median of `--iters` launches after 3 warm-ups.  achieved GB/s uses the
ALGORITHMIC bytes (each input element read once, each output element written once); the fraction
is against MEASURED_PEAKS.json hbm_gbs.  One JSON object per line on stdout, each carrying the
nvidia-smi clock / throttle-reason samples taken while that row was being measured (a background
sampler with timestamps; rows are printed when the sweep ends).

    python tools/sweep.py [--iters 10] [--only cfg3]
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_stereo_matcher_b200 as rsm  # noqa: E402

DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FMA pipe at the boost clock the sweeps record
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


class TimedClockSampler:
    """nvidia-smi sampled every 50 ms with timestamps for the whole sweep; window(t0, t1) -> the clocks record of
    the samples taken in [t0 - 0.1 s, t1 + 0.1 s]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc, self.path, self.samples = None, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", "0"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        self.samples = []
        if not self.proc:
            return
        time.sleep(0.2)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        for line in open(self.path).read().strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 8:
                continue
            try:
                t = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.samples.append((t, float(f[1]), float(f[2]), float(f[3]), [n for n, v in zip(
                    ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)

    def window(self, t0, t1):
        sel = [s for s in self.samples if t0 - 0.1 <= s[0] <= t1 + 0.1]
        if not sel:
            return {"sm_mhz": None, "samples": 0, "reasons": []}
        return {"sm_mhz": statistics.median(s[1] for s in sel), "sm_max_mhz": sel[0][2], "power_w_max": max(s[3] for s in sel),
                "reasons": sorted({r for s in sel for r in s[4]}), "samples": len(sel)}


_ROWS = []
_T0 = [time.time()]


def report(cfg, op, dtype, shape, ms, nbytes, flops=0):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rec = {"cfg": cfg, "op": op, "dtype": dtype, "shape": shape, "ms": round(ms, 5), "algorithmic_bytes": nbytes,
           "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 3)}
    if flops:
        rec["TFLOPs"] = round(flops / (ms * 1e-3) / 1e12, 2)
        if dtype == "f32":      # SIMT kernels: 148 SMs x 128 FMA lanes x 2 flops x 1.965 GHz = 74.4 TFLOP/s
            rec["frac_of_fp32_fma_peak"] = round(flops / (ms * 1e-3) / 1e12 / FFMA_PEAK_TFLOPS, 3)
    now = time.time()
    _ROWS.append((rec, _T0[0], now))      # measured between the previous report and this one
    _T0[0] = now
    print(f"  .. {cfg} {op} {dtype} {rec['ms']} ms {rec['frac_of_measured_hbm']}", file=sys.stderr, flush=True)


def sustained(cfg, op, dtype, shape, fn, nbytes, seconds=2.0):
    """>= `seconds` of back-to-back launches (no L2 flush: the working sets are >> L2): does the fraction hold under load?"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    reps = max(20, int(seconds / (a.elapsed_time(b) * 1e-3)))
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    report(cfg, op + f" [sustained {a.elapsed_time(b) * 1e-3:.1f} s, {reps} launches back to back]", dtype, shape, ms, nbytes)


def feats(n, c, h, w, dt):
    g = torch.Generator(device="cuda").manual_seed(1234)
    return (torch.randn((n, c, h, w), device="cuda", generator=g).to(DT[dt]),
            torch.randn((n, c, h, w), device="cuda", generator=g).to(DT[dt]))


def sweep_volumes(cfg, n, c, h, w, d, g, dtypes, iters, ops):
    for dt in dtypes:
        e = 2 if dt != "f32" else 4
        L, R = feats(n, c, h, w, dt)
        fin = 2 * n * c * h * w * e
        shp = dict(N=n, C=c, H=h, W=w, D=d, G=g)
        with torch.no_grad():
            if "concat" in ops:
                report(cfg, "concat_fwd", dt, shp, timed(lambda: rsm.concat_volume(L, R, d), iters), fin + 2 * n * c * h * w * d * e)
            if "interweave" in ops:
                report(cfg, "interweave_fwd", dt, shp, timed(lambda: rsm.interweave(L, R), iters), 2 * fin)
            if "difference" in ops:
                report(cfg, "difference_fwd", dt, shp, timed(lambda: rsm.difference_volume(L, R, d), iters), fin + n * c * d * h * w * e)
            if "groupwise" in ops:
                report(cfg, "groupwise_fwd", dt, shp, timed(lambda: rsm.groupwise_volume(L, R, g, d), iters),
                       fin + n * g * h * w * d * e, 2 * n * c * h * w * d)
            if "inner" in ops:
                report(cfg, "inner_mean_fwd", dt, shp, timed(lambda: rsm.inner_product_volume(L, R, d, mean=True), iters),
                       fin + n * d * h * w * e, 2 * n * c * h * w * d)
            if "fused" in ops and d <= 512:
                report(cfg, "inner_regress_fused_fwd", dt, shp, timed(lambda: rsm.inner_product_regress(L, R, d), iters),
                       fin + n * h * w * 20, 2 * n * c * h * w * d)
        if "bwd" in ops:
            Lg, Rg = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
            for name, fn, nb, fl in (
                ("concat_bwd", lambda: rsm.concat_volume(Lg, Rg, d), 2 * n * c * h * w * d * e + fin, 0),
                ("groupwise_bwd", lambda: rsm.groupwise_volume(Lg, Rg, g, d), n * g * h * w * d * e + 2 * fin, 4 * n * c * h * w * d),
                ("inner_mean_bwd", lambda: rsm.inner_product_volume(Lg, Rg, d, mean=True), n * d * h * w * e + 2 * fin, 4 * n * c * h * w * d),
            ):
                if name.split("_")[0] not in ops:
                    continue
                out = fn()
                go = torch.randn_like(out)
                report(cfg, name, dt, shp, timed(lambda: torch.autograd.grad(out, (Lg, Rg), go, retain_graph=True), iters), nb, fl)
                del out, go
        del L, R
        torch.cuda.empty_cache()


def sweep_regress(cfg, n, d, h, w, dtypes, iters):
    for dt in dtypes:
        e = 2 if dt != "f32" else 4
        g = torch.Generator(device="cuda").manual_seed(1234)
        cost = (torch.randn((n, d, h, w), device="cuda", generator=g) * 4).to(DT[dt])
        shp = dict(N=n, D=d, H=h, W=w)
        with torch.no_grad():
            report(cfg, "regress_fwd(soft+argmin+argmax)", dt, shp, timed(lambda: rsm.regress(cost), iters),
                   n * d * h * w * e + n * h * w * (e + 16))
            report(cfg, "soft_argmax_fwd", dt, shp, timed(lambda: rsm.soft_argmax(cost), iters), n * d * h * w * e + n * h * w * e)
            report(cfg, "hard_argmin_fwd", dt, shp, timed(lambda: rsm.hard_argmin(cost), iters), n * d * h * w * e + n * h * w * 8)
        cg = cost.requires_grad_(True)
        out = rsm.soft_argmax(cg)
        go = torch.randn_like(out)
        report(cfg, "soft_argmax_bwd", dt, shp, timed(lambda: torch.autograd.grad(out, cg, go, retain_graph=True), iters),
               2 * n * d * h * w * e + 3 * n * h * w * 4)
        del cost, cg, out, go
        torch.cuda.empty_cache()


def sweep_tail(cfg, b, dc, hc, wc, iters):
    g = torch.Generator(device="cuda").manual_seed(1234)
    cost = torch.randn((b, dc, hc, wc), device="cuda", generator=g) * 3
    d, h, w = 4 * dc, 4 * hc, 4 * wc
    shp = dict(B=b, Dc=dc, Hc=hc, Wc=wc, D=d, H=h, W=w)
    nb = b * dc * hc * wc * 4 + b * h * w * 4
    with torch.no_grad():
        ms = timed(lambda: rsm.v4_head(cost, d, h, w), iters)
        report(cfg, "v4_head_fwd", "f32", {**shp, "Gexp_per_s": round(b * d * h * w / (ms * 1e-3) / 1e9, 1)}, ms, nb)
        report(cfg, "v4_head_fwd+argmin+argmax", "f32", shp,
               timed(lambda: rsm.upsample_regress(cost, d, h, w, argmin=True, argmax=True), iters), nb + b * h * w * 16)
        # the unfused 4-kernel path this replaces (stock torch on the same device), for the speed-up figure
        def unfused():
            fine = torch.nn.functional.interpolate(cost.unsqueeze(1), [d, h, w], mode="trilinear").squeeze(1)
            p = torch.softmax(fine, 1)
            return (p * torch.arange(d, device="cuda", dtype=p.dtype).view(1, -1, 1, 1)).sum(1)
        if b * d * h * w * 4 * 3 < 60e9:
            report(cfg, "v4_head_unfused_torch(reference op sequence on GPU)", "f32", shp, timed(unfused, max(3, iters // 3)), nb)
    cg = cost.requires_grad_(True)
    out = rsm.v4_head(cg, d, h, w)
    go = torch.randn_like(out)
    # algorithmic bytes: cost read, its gradient written, incoming gradient + expectation + lse read (the tile-partial
    # workspace of the x4 head, 40 floats per tile and slice, is not algorithmic); the op is issue-bound like the forward
    ms = timed(lambda: torch.autograd.grad(out, cg, go, retain_graph=True), iters)
    report(cfg, "v4_head_bwd", "f32", {**shp, "Gexp_per_s": round(b * d * h * w / (ms * 1e-3) / 1e9, 1)}, ms,
           2 * b * dc * hc * wc * 4 + 3 * b * h * w * 4)


def sweep_v4_volume(cfg, iters):
    """MobileStereoNetV4's whole per-disparity volume (48 x (interweave -> 3 Conv3d -> 1x1)) as the fused op,
    with the reference's own module weights when baseline/_ref is installed; flops = the reference loop's 2*MACs."""
    try:
        from oracle import ref_loader
        ref = ref_loader.load()
    except Exception:
        return
    torch.manual_seed(1234)
    net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).cuda().eval()
    for b in (1, 8):
        for dt in ("f32", "f16"):
            L, R = feats(b, 32, 96, 312, dt)
            px = b * 96 * 312 * 48
            flops = 2.0 * px * (8 * 72 * 16 + 2 * 576 * 32 + 576 * 16 + 16)
            e = 2 if dt != "f32" else 4
            with torch.no_grad():
                report(cfg, "v4_cost_volume_fused (K1 maps + 2 tcgen05 implicit GEMMs)", dt, dict(B=b, C=32, H=96, W=312, D=48),
                       timed(lambda: rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, 48), iters),
                       2 * b * 32 * 96 * 312 * e + px * e, flops)
    L, R = feats(1, 32, 96, 312, "f32")
    v4 = ref.v4

    def loop():
        vol = L.new_zeros([1, 48, 96, 312])
        for i in range(48):
            x = v4.interweave_tensors(L[:, :, :, i:], R[:, :, :, : 312 - i])
            vol[:, i, :, i:] = net.volume11(torch.squeeze(net.conv3d(torch.unsqueeze(x, 1)), 2))[:, 0]
        return vol
    with torch.no_grad():
        report(cfg, "v4_cost_volume_reference_loop (cuDNN, TF32 allowed; reference op sequence on GPU)", "f32",
               dict(B=1, C=32, H=96, W=312, D=48), timed(loop, 3), 2 * 32 * 96 * 312 * 4 + 96 * 312 * 48 * 4,
               2.0 * 96 * 312 * 48 * (8 * 72 * 16 + 2 * 576 * 32 + 576 * 16 + 16))


def sweep_warp(cfg, n, c, h, w, iters):
    """refinement warp (8f-2) with a smooth disparity field (what RefineNet feeds it) and with white-noise flow"""
    g = torch.Generator(device="cuda").manual_seed(1234)
    image = torch.randn((n, c, h, w), device="cuda", generator=g)
    yy, xx = torch.meshgrid(torch.arange(h, device="cuda", dtype=torch.float32),
                            torch.arange(w, device="cuda", dtype=torch.float32), indexing="ij")
    smooth = (20 + 12 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0)).expand(n, 1, h, w).contiguous()
    noise = torch.rand((n, 1, h, w), device="cuda", generator=g) * 40
    shp = dict(N=n, C=c, H=h, W=w)
    nb = (2 * n * c * h * w + n * h * w) * 4
    for tag, flow in (("smooth", smooth), ("noise", noise)):
        with torch.no_grad():
            report(cfg, f"warp_fwd({tag} flow)", "f32", shp, timed(lambda: rsm.warp_by_flow_map(image, flow), iters), nb)
        im, fl = image.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        out = rsm.warp_by_flow_map(im, fl)
        go = torch.randn_like(out)
        report(cfg, f"warp_bwd({tag} flow)", "f32", shp,
               timed(lambda: torch.autograd.grad(out, (im, fl), go, retain_graph=True), iters),
               (4 * n * c * h * w + 2 * n * h * w) * 4)
    def ref():
        grid_x = (xx.view(1, 1, h, w) - smooth[:, 0].view(n, 1, h, w)).permute(0, 2, 3, 1)
        grid_y = yy.view(1, h, w, 1).repeat(n, 1, 1, 1)
        grid = torch.cat((2.0 * grid_x / (w - 1.0) - 1.0, 2.0 * grid_y / (h - 1.0) - 1.0), dim=-1)
        return torch.nn.functional.grid_sample(image, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    with torch.no_grad():
        report(cfg, "warp_fwd_torch(reference op sequence on GPU, smooth flow)", "f32", shp, timed(ref, iters), nb)


def sweep_prepost(cfg, n, h, w, align, factor, iters):
    """pre / post steps (8f-3): normalise + pad of an RGB frame; scale + resize + crop + negate of a 1/factor map"""
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(1234)
    img = torch.rand((n, 3, h, w), device="cuda", generator=g) * 255
    hp, wp = h + (align - h % align) % align, w + (align - w % align) % align
    shp = dict(N=n, H=h, W=w, Hp=hp, Wp=wp)
    nb = n * 3 * (h * w + hp * wp) * 4
    with torch.no_grad():
        report(cfg, "prepare_input", "f32", shp, timed(lambda: rsm.prepare_input(img, align), iters), nb)
        report(cfg, "prepare_input_torch(reference op sequence on GPU)", "f32", shp,
               timed(lambda: F.pad((2.0 * (img / 255.0) - 1.0).contiguous(), (0, wp - w, 0, hp - h)), iters), nb)
    disp = torch.rand((n, 1, hp // factor, wp // factor), device="cuda", generator=g) * 20
    nb = n * (disp.shape[2] * disp.shape[3] + h * w) * 4
    for mode in ("nearest", "bilinear"):
        kw = {} if mode == "nearest" else dict(mode="bilinear", align_corners=False)
        with torch.no_grad():
            report(cfg, f"finalize_disparity({mode}, x{factor})", "f32", shp,
                   timed(lambda: rsm.finalize_disparity(disp, (hp, wp), (h, w), mode=mode), iters), nb)
            report(cfg, f"finalize_torch({mode}, x{factor}; reference op sequence on GPU)", "f32", shp,
                   timed(lambda: -1.0 * F.interpolate(disp * (float(wp) / disp.shape[3]), (hp, wp), **kw)[:, :, :h, :w], iters), nb)
        dg = disp.clone().requires_grad_(True)
        out = rsm.finalize_disparity(dg, (hp, wp), (h, w), mode=mode)
        go = torch.randn_like(out)
        report(cfg, f"finalize_disparity_bwd({mode}, x{factor})", "f32", shp,
               timed(lambda: torch.autograd.grad(out, dg, go, retain_graph=True), iters), nb)


def sweep_loss(cfg, n, h, w, sizes, iters):
    """SequenceLoss + metrics (8f-4) for a training batch: forward + backward of the loss, and the metrics, against
    the reference's op sequence (torch on the same GPU; its .item() / assert synchronisations left out)"""
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(1234)
    gt = torch.randn((n, 1, h, w), device="cuda", generator=g) * 40
    valid = (torch.rand((n, h, w), device="cuda", generator=g) > 0.2).float()
    preds = [(torch.randn((n, 1) + s, device="cuda", generator=g) * 10).requires_grad_(True) for s in sizes]
    shp = dict(N=n, H=h, W=w, preds=[list(s) for s in sizes])
    nb = (sum(p.numel() for p in preds) * 2 + len(preds) * 2 * n * h * w * 2) * 4
    loss_fn = rsm.SequenceLoss(0.9, 700, check_finite=False)

    def ours():
        return torch.autograd.grad(loss_fn(preds, gt, valid), preds)

    def ref():
        mask = ((valid >= 0.5) & (torch.sum(gt ** 2, dim=1).sqrt() < 700)).unsqueeze(1)
        total = 0.0
        for i, p in enumerate(preds):
            q = p
            if q.shape != gt.shape:
                q = F.interpolate(q * (float(w) / q.shape[-1]), (h, w))
            el = F.smooth_l1_loss(gt, q, reduction="none", beta=1.0) if i == len(preds) - 1 else F.l1_loss(gt, q, reduction="none")
            total = total + 0.9 ** (len(preds) - 1 - i) * el[mask].mean()
        return torch.autograd.grad(total, preds)

    report(cfg, "sequence_loss fwd+bwd", "f32", shp, timed(ours, iters), nb)
    report(cfg, "sequence_loss_torch fwd+bwd (reference op sequence on GPU, asserts left out)", "f32", shp, timed(ref, max(3, iters // 3)), nb)
    last = preds[-1].detach()
    if last.shape == gt.shape:
        report(cfg, "flow_map_metrics (device vector, no sync)", "f32", shp, timed(lambda: rsm.flow_map_metrics(gt, last, valid), iters), 3 * n * h * w * 4)

        def ref_metrics():
            epe = torch.sum((last - gt) ** 2, dim=1).sqrt().view(-1)[(valid >= 0.5).view(-1)]
            return [epe.mean(), (epe < 0.5).float().mean(), (epe < 1).float().mean(), (epe < 3).float().mean(),
                    (epe < 5).float().mean(), torch.min(last[0]), torch.max(last[0])]
        report(cfg, "flow_map_metrics_torch (reference op sequence on GPU, .item() left out)", "f32", shp, timed(ref_metrics, iters), 3 * n * h * w * 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    want = lambda c: not a.only or c in a.only.split(",")
    rsm.load_library()
    sampler = TimedClockSampler()
    sampler.start()
    _T0[0] = time.time()
    if want("cfg1"):
        sweep_volumes("cfg1", 1, 32, 48, 156, 24, 8, ["f32"], a.iters, {"difference", "bwd"})
        sweep_volumes("cfg1", 8, 32, 48, 156, 24, 8, ["f32", "bf16"], a.iters, {"difference", "bwd"})   # training batch of 8
        sweep_regress("cfg1", 1, 24, 48, 156, ["f32"], a.iters)
    if want("cfg2"):
        for c in (16, 64):
            sweep_volumes("cfg2", 32, c, 144, 240, 48, 1, ["f32", "bf16"], a.iters, {"inner", "fused", "bwd"})
    if want("cfg3"):
        sweep_volumes("cfg3", 8, 32, 96, 312, 48, 8, ["f32", "bf16"], a.iters, {"concat", "groupwise", "interweave", "bwd"})
        sweep_tail("cfg3", 8, 48, 96, 312, a.iters)
        sweep_warp("cfg3", 8, 32, 96, 312, a.iters)      # v3 RefineNet: 32-channel features at 1/4 resolution
        sweep_warp("cfg3", 8, 3, 192, 624, a.iters)      # v2 RefineNet: RGB at 1/2 resolution
        sweep_regress("cfg3", 8, 192, 384, 1248, ["f32"], a.iters)
        sweep_prepost("cfg3", 8, 375, 1242, 64, 4, a.iters)   # raw KITTI frame -> 384 x 1280, 1/4-resolution map back
        sweep_prepost("cfg3", 8, 384, 1248, 8, 8, a.iters)
        sweep_loss("cfg3", 8, 384, 1248, [(96, 312), (192, 624), (384, 1248)], a.iters)
        sweep_loss("cfg3", 8, 384, 1248, [(384, 1248)] * 6, a.iters)      # DispNetC: six full-size predictions
        sweep_v4_volume("cfg3", a.iters)
        L, R = feats(8, 32, 96, 312, "f32")
        with torch.no_grad():
            sustained("cfg3", "concat_fwd", "f32", dict(N=8, C=32, H=96, W=312, D=48), lambda: rsm.concat_volume(L, R, 48),
                      2 * 8 * 32 * 96 * 312 * 4 * (1 + 48))
        del L, R
    if want("cfg4"):
        # the full (C, G, D) grid, forward and backward: concat / inner depend on (C, D) only and are run once per pair
        seen = set()
        for c in (32, 64, 128):
            for d in (48, 96, 192):
                for g in (8, 16, 32):
                    ops = {"groupwise", "bwd"}
                    if (c, d) not in seen:
                        ops |= {"concat", "inner", "fused"}
                        seen.add((c, d))
                    sweep_volumes("cfg4", 1, c, 270, 480, d, g, ["f32", "bf16"], a.iters, ops)
    if want("cfg5"):
        cost = torch.randn((4, 192, 1080, 1920), device="cuda") * 4
        with torch.no_grad():
            sustained("cfg5", "soft_argmax_fwd", "f32", dict(N=4, D=192, H=1080, W=1920), lambda: rsm.soft_argmax(cost),
                      4 * 192 * 1080 * 1920 * 4 + 4 * 1080 * 1920 * 4)
        del cost
        torch.cuda.empty_cache()
        sweep_regress("cfg5", 1, 192, 1080, 1920, ["f32", "bf16"], a.iters)
        sweep_regress("cfg5", 4, 192, 1080, 1920, ["f32"], a.iters)
        sweep_tail("cfg5", 1, 48, 270, 480, a.iters)
    sampler.stop()
    for rec, t0, t1 in _ROWS:
        rec["clocks"] = sampler.window(t0, t1)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
