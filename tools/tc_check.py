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
print("--- fused regress (tcgen05)")
for (n,c,h,w,d) in [(1,16,2,128,16),(1,16,3,240,48),(2,64,5,240,48),(1,32,4,312,48),(1,64,2,480,128),(1,16,2,67,19),(1,128,2,480,192),(1,32,3,200,130),(2,16,2,300,260)]:
    for dn,dt in (("bf16",torch.bfloat16),):
        rng=np.random.default_rng(1)
        l=round_to(rng.standard_normal((n,c,h,w)).astype(np.float32)*0.5,dn); r=round_to(rng.standard_normal((n,c,h,w)).astype(np.float32)*0.5,dn)
        L=torch.from_numpy(l).cuda().to(dt); R=torch.from_numpy(r).cuda().to(dt)
        so,mi,ma=rsm.inner_product_regress(L,R,d)
        vol=oracle.inner_product_volume(l,r,d,out_dtype=np.float32)
        es=np.abs(so.cpu().numpy()-oracle.soft_argmax(vol)).max()
        mm=(mi.cpu().numpy()!=oracle.hard_argmin(vol)).mean(); mx=(ma.cpu().numpy()!=oracle.hard_argmax(vol)).mean()
        print((n,c,h,w,d),dn,"soft err",es,"argmin mismatch",mm,"argmax mismatch",mx, flush=True)
# dyadic: bit exact argmin/argmax expected
rng=np.random.default_rng(7); n,c,h,w,d=2,32,6,156,24
l=(rng.integers(-8,9,(n,c,h,w))/8.0).astype(np.float32); r=(rng.integers(-8,9,(n,c,h,w))/8.0).astype(np.float32)
L=torch.from_numpy(l).cuda().bfloat16(); R=torch.from_numpy(r).cuda().bfloat16()
so,mi,ma=rsm.inner_product_regress(L,R,d); vol=oracle.inner_product_volume(l,r,d)
print("dyadic argmin equal", np.array_equal(mi.cpu().numpy(),oracle.hard_argmin(vol)), "argmax equal", np.array_equal(ma.cpu().numpy(),oracle.hard_argmax(vol)),
      "volume equal", np.array_equal(rsm.inner_product_volume(L,R,d,out_dtype=torch.float32).cpu().numpy(), vol))
