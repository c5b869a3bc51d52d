import sys, torch, os
sys.path.insert(0,'.')
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["RSM_TC_DBGBUF"] = str(buf.data_ptr())
import realtime_stereo_matcher_b200 as rsm
for name, (n, c, h, w, d) in {"cfg2 C64 D48": (32, 64, 144, 240, 48), "cfg4 C128 D192": (1, 128, 270, 480, 192)}.items():
    L = torch.randn(n, c, h, w, device="cuda").bfloat16(); R = torch.randn(n, c, h, w, device="cuda").bfloat16()
    for mode in ("volume", "fused"):
        for _ in range(3):
            if mode == "volume": rsm.inner_product_volume(L, R, d, mean=True)
            else: rsm.inner_product_regress(L, R, d, mean=True)
        torch.cuda.synchronize()
        b = buf.cpu().numpy()[:60].reshape(12, 5)
        print(name, mode)
        for wv in range(12):
            if b[wv,4]: print("  warp %2d: tiles %d  per tile cycles: wait %.0f copy %.0f math %.0f rest/combine %.0f" % (wv, b[wv,4], b[wv,0]/b[wv,4], b[wv,1]/b[wv,4], b[wv,2]/b[wv,4], b[wv,3]/b[wv,4]))
        buf.zero_()
