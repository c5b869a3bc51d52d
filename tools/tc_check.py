import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle, realtime_stereo_matcher_b200 as rsm
from golden_io import round_to
torch.manual_seed(0)
for (n,c,h,w,d) in [(1,16,2,128,16),(1,16,3,240,48),(2,64,5,240,48),(1,32,4,312,48),(1,128,2,480,192),(1,16,2,67,19)]:
    for dn,dt in (("bf16",torch.bfloat16),("fp16",torch.float16)):
        rng=np.random.default_rng(1)
        l=round_to(rng.standard_normal((n,c,h,w)).astype(np.float32),dn); r=round_to(rng.standard_normal((n,c,h,w)).astype(np.float32),dn)
        L=torch.from_numpy(l).cuda().to(dt); R=torch.from_numpy(r).cuda().to(dt)
        out=rsm.inner_product_volume(L,R,d,out_dtype=torch.float32).float().cpu().numpy()
        ref=oracle.inner_product_volume(l,r,d,out_dtype=np.float32)
        err=np.abs(out-ref).max(); nan=np.isnan(out).sum()
        print((n,c,h,w,d),dn,"maxerr",err,"nan",nan,"ref max",np.abs(ref).max(), flush=True)
