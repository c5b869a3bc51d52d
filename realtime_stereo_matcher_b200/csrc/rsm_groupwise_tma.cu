// Group-wise correlation volume, TMA-fed persistent form: TorchGroupwiseCost.forward (cost_volume/groupwise.py:12-56)
//     V[n, g, y, x, d] = (1 / cpg) * sum_{c in group g} L[n, c, y, x] * R[n, c, y, x - d]   (x >= d, else 0)
// for groups of <= 32 channels and feature rows TMA can address.  The op writes D values per input pixel and group
// (368 MB for 61 MB of features at BASELINE config 3): it is bound by HBM writes -- if the stores never wait.  The
// first kernel (corr_fwd_kernel, rsm_corr.cu) stages a slab with thousands of 4-byte cp.async per CTA, waits, computes
// and stores: with 4-5 such CTAs per SM it reaches 0.75 of the HBM roofline in fp32 and takes the SAME time with 16-bit
// tensors (0.41) -- it is bound by the exposed staging latency, not by bytes.  Here one persistent CTA per SM runs a
// ring of operand stages: a producer lane issues two TMA box loads per tile (left segment 64 pixels x CT channels,
// right window (64 + Dch) pixels x CT channels, zero-filled outside the image, raw dtype) up to five tiles ahead of
// twelve consumer warps, which turn a landed stage into 4(x) x 8(d) register tiles per group (two or three LDS.128 /
// LDS.64 per channel, 16-bit features widened in registers) and write them with 256-bit (fp32) / 128-bit (16-bit)
// streaming stores, 192 contiguous bytes per pixel.  Loads, FMAs and stores of different tiles overlap by
// construction; no thread ever waits for a load it issued itself.
// Measured (B200, config 3: 8 x 32 x 96 x 312, G = 8, D = 48): 16-bit 82 -> 68 us; fp32 stays at 86 us = 5.0 TB/s, which
// is where every variant of this op lands -- contiguous 1 KB per warp store, staging tiles + TMA stores, folded
// scaling (-30 % instructions) all measured 85-86 us while a plain fill of the same 368 MB takes 55 us; ncu shows no
// unit above 60 % (issue 52 %, LSU wavefronts 60 %, L1->XBAR writes 51 %, DRAM 58 %) at 12.7 resident warps per SM:
// the 4 x 8 register tile costs 128 registers, and the op is bound by the parallelism that leaves (24 consumer warps
// at 80 registers spill and measured 10-50 % slower).  Config 4 points with >= 8 channels per group gain 1.1-1.4x.
#include <cuda.h>

#include <type_traits>

#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int GT_TX = 64;            // pixels per tile
constexpr int GT_XT = 4, GT_DT = 8;  // register tile: 4 pixels x 8 disparities
constexpr int GT_NTX = GT_TX / GT_XT;
constexpr int GT_CT = 32;            // channels per tile (upper bound)
constexpr int GT_DCH = 64;           // disparities per tile (upper bound)
constexpr int GT_MAXSTAGES = 6;
constexpr int GT_CONS_WARPS = 12;
constexpr int GT_THREADS = 32 * (GT_CONS_WARPS + 1);

struct GtGeom {
  int C, G, cpg, H, W, D;
  int dch, ndch, ntd;     // disparities per tile (multiple of 8), tiles along D, thread tiles along D
  int gt, gblocks;        // groups per tile, tiles along the groups
  int xtiles;
  int rw;                 // right window width = 64 + dch
  int l_bytes, tx_bytes, stage_bytes, nstage;   // left part, both parts, stage pitch (128-byte multiple), ring depth
  int nsets;              // consumer sets of 16 * ntd threads (each takes every nsets-th group of a tile)
  int pow2;
  int vec;                // D % 8 == 0 and the output is 8-element aligned: whole-vector stores
  int64_t tiles;
};

struct GtTile {
  int n, y, gb, dc, xt;
  __device__ __forceinline__ void advance(const GtGeom& g) {
    if (++xt < g.xtiles) return;
    xt = 0;
    if (++dc < g.ndch) return;
    dc = 0;
    if (++gb < g.gblocks) return;
    gb = 0;
    if (++y < g.H) return;
    y = 0; ++n;
  }
};

template <typename T> __device__ __forceinline__ void gt_load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void gt_load4<float>(const float* p, float (&v)[4]) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <> __device__ __forceinline__ void gt_load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void gt_load4<__half>(const __half* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ void gt_store8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void gt_store8(__half* p, const float* v) {
  union { uint4 u; __half2 h[4]; } t;
#pragma unroll
  for (int k = 0; k < 4; ++k) t.h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
  __stcs(reinterpret_cast<uint4*>(p), t.u);
}
__device__ __forceinline__ void gt_store8(__nv_bfloat16* p, const float* v) {
  union { uint4 u; __nv_bfloat162 h[4]; } t;
#pragma unroll
  for (int k = 0; k < 4; ++k) t.h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  __stcs(reinterpret_cast<uint4*>(p), t.u);
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(GT_THREADS, 1)
groupwise_tma_kernel(Tout* __restrict__ out, GtGeom g, const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.nstage * g.stage_bytes);
  const uint32_t full = smem_u32(bars), empty = full + 8 * GT_MAXSTAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_MAXSTAGES; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, GT_CONS_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t per = (g.tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_beg = min((int64_t)blockIdx.x * per, g.tiles), t_end = min(t_beg + per, g.tiles);
  const uint32_t ntl = (uint32_t)(t_end - t_beg);
  GtTile first;
  {
    uint32_t t = (uint32_t)t_beg;
    first.xt = (int)(t % (uint32_t)g.xtiles); t /= (uint32_t)g.xtiles;
    first.dc = (int)(t % (uint32_t)g.ndch); t /= (uint32_t)g.ndch;
    first.gb = (int)(t % (uint32_t)g.gblocks); t /= (uint32_t)g.gblocks;
    first.y = (int)(t % (uint32_t)g.H);
    first.n = (int)(t / (uint32_t)g.H);
  }
  const uint32_t nst = (uint32_t)g.nstage;

  if (warp == GT_CONS_WARPS) {
    // ================================================================ TMA producer (one lane)
    if (lane == 0) {
      uint32_t s = 0, p = 0;
      GtTile tc = first;
      for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g)) {
        const int x0 = tc.xt * GT_TX, c0 = tc.gb * g.gt * g.cpg, dc0 = tc.dc * g.dch;
        mbar_wait(empty + 8 * s, p ^ 1);
        mbar_expect_tx(full + 8 * s, (uint32_t)g.tx_bytes);
        const uint32_t dst = smem_u32(smem) + s * (uint32_t)g.stage_bytes;
        tma_load_4d(dst, &tmL, full + 8 * s, x0, tc.y, c0, tc.n);
        tma_load_4d(dst + (uint32_t)g.l_bytes, &tmR, full + 8 * s, x0 - dc0 - g.dch, tc.y, c0, tc.n);
        if (++s == nst) { s = 0; p ^= 1; }
      }
    }
  } else {
    // ================================================================ consumers
    // thread -> (set, pixel quad tx, disparity octet td), td fastest: a warp writes whole 192-byte pixel runs
    const int per_set = GT_NTX * g.ntd;
    const int set = threadIdx.x / per_set, idx = threadIdx.x - set * per_set;
    const int td = idx % g.ntd, tx = idx / g.ntd;
    const bool active = set < g.nsets;
    const int wstart = g.dch + GT_XT * tx - GT_DT * td - GT_DT;     // first of the 12 window values of this thread
    const float inv = 1.f / (float)g.cpg, cnt = (float)g.cpg;
    uint32_t s = 0, p = 0;
    GtTile tc = first;
    for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g)) {
      mbar_wait(full + 8 * s, p);
      if (active) {
        const Tin* sL = reinterpret_cast<const Tin*>(smem + (size_t)s * g.stage_bytes);
        const Tin* sR = reinterpret_cast<const Tin*>(smem + (size_t)s * g.stage_bytes + g.l_bytes);
        const int xb = tc.xt * GT_TX + GT_XT * tx, db = tc.dc * g.dch + GT_DT * td;
        const int g0 = tc.gb * g.gt, g1 = min(g.G, g0 + g.gt);
        for (int grp = g0 + set; grp < g1; grp += g.nsets) {
          float acc[GT_XT][GT_DT];
          const Tin* pl = sL + (grp - g0) * g.cpg * GT_TX + GT_XT * tx;
          const Tin* pw = sR + (grp - g0) * g.cpg * g.rw + wstart;
          // the mean over the group's channels: a power-of-two count scales the left values (exact, so identical to
          // scaling the sum); any other count divides the finished sum like the reference does
          const float lscale = g.pow2 ? inv : 1.f;
          auto channel = [&](int c, bool first_ch) {
            float l[4], w[12];
            gt_load4<Tin>(pl + c * GT_TX, l);
            gt_load4<Tin>(pw + c * g.rw, *reinterpret_cast<float(*)[4]>(&w[0]));
            gt_load4<Tin>(pw + c * g.rw + 4, *reinterpret_cast<float(*)[4]>(&w[4]));
            gt_load4<Tin>(pw + c * g.rw + 8, *reinterpret_cast<float(*)[4]>(&w[8]));
#pragma unroll
            for (int i = 0; i < GT_XT; ++i) {
              const float li = l[i] * lscale;
#pragma unroll
              for (int j = 0; j < GT_DT; ++j) acc[i][j] = first_ch ? li * w[8 + i - j] : fmaf(li, w[8 + i - j], acc[i][j]);
            }
          };
          channel(0, true);
#pragma unroll 3
          for (int c = 1; c < g.cpg; ++c) channel(c, false);
          if (!g.pow2) {
#pragma unroll
            for (int i = 0; i < GT_XT; ++i)
#pragma unroll
              for (int j = 0; j < GT_DT; ++j) acc[i][j] = acc[i][j] / cnt;
          }
          if (xb < db + GT_DT - 1) {
#pragma unroll
            for (int i = 0; i < GT_XT; ++i)
#pragma unroll
              for (int j = 0; j < GT_DT; ++j)
                if (xb + i < db + j) acc[i][j] = 0.f;
          }
          if (db < g.D) {
            Tout* o = out + ((((int64_t)tc.n * g.G + grp) * g.H + tc.y) * g.W + xb) * g.D + db;
#pragma unroll
            for (int i = 0; i < GT_XT; ++i, o += g.D) {
              if (xb + i >= g.W) break;
              if (g.vec) gt_store8(o, &acc[i][0]);
              else {
#pragma unroll
                for (int j = 0; j < GT_DT; ++j)
                  if (db + j < g.D) o[j] = from_f<Tout>(acc[i][j]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * s);                 // this warp is done reading the stage
      if (++s == nst) { s = 0; p ^= 1; }
    }
  }
}

static bool gt_tmap(CUtensorMap* m, const rsm_feat& f, int es, CUtensorMapDataType dt, int64_t W, int64_t H, int64_t C, int64_t N,
                    uint32_t bw, uint32_t bc) {
  const TmapEncodeFn enc = tmap_encoder();
  if (!enc || f.stride_w != 1 || !aligned_to(f.data, 16)) return false;
  const int64_t st[3] = {f.stride_h * es, f.stride_c * es, f.stride_n * es}, ext[3] = {H, C, N};
  cuuint64_t gstr[3];
  for (int i = 0; i < 3; ++i) {
    int64_t v = st[i];
    if (ext[i] == 1 && (v % 16 != 0 || v <= 0)) v = 16;
    if (v <= 0 || v % 16 != 0 || v >= (1LL << 40)) return false;
    gstr[i] = (cuuint64_t)v;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  const cuuint32_t estr[4] = {1, 1, 1, 1}, box[4] = {bw, 1, bc, 1};
  return enc(m, dt, 4, const_cast<void*>(f.data), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename Tin, typename Tout>
static int gt_launch(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, GtGeom g, cudaStream_t st) {
  const int es = (int)sizeof(Tin);
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                         : (std::is_same<Tin, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  alignas(64) CUtensorMap tmL, tmR;
  memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
  const uint32_t bc = (uint32_t)(g.gt * g.cpg);
  if (!gt_tmap(&tmL, left, es, dt, g.W, g.H, g.C, N, GT_TX, bc) || !gt_tmap(&tmR, right, es, dt, g.W, g.H, g.C, N, (uint32_t)g.rw, bc))
    return RSM_ERR_UNSUPPORTED_CONFIG;
  g.l_bytes = (int)bc * GT_TX * es;
  g.tx_bytes = (int)bc * (GT_TX + g.rw) * es;
  g.stage_bytes = (g.tx_bytes + 127) / 128 * 128;          // TMA destinations: 128-byte aligned (the left part always is)
  const size_t extra = 256 + 128;
  g.nstage = GT_MAXSTAGES;
  while (g.nstage > 2 && (size_t)g.nstage * g.stage_bytes + extra > 200 * 1024) --g.nstage;
  const size_t smem = (size_t)g.nstage * g.stage_bytes + extra;
  if (smem > 227 * 1024) return RSM_ERR_UNSUPPORTED_CONFIG;
  auto k = groupwise_tma_kernel<Tin, Tout>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return finish_launch("rsm_groupwise_fwd(tma)");
  const unsigned grid = (unsigned)(g.tiles < kNumSMs ? g.tiles : kNumSMs);
  k<<<grid, GT_THREADS, smem, st>>>((Tout*)out, g, tmL, tmR);
  return finish_launch("rsm_groupwise_fwd(tma)");
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when this form does not apply (the caller falls back to corr_fwd_kernel)
int launch_groupwise_tma(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                         int64_t D, int64_t G, int in_dtype, int out_dtype, cudaStream_t st) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0 || C % G != 0) return RSM_ERR_UNSUPPORTED_CONFIG;
  const int64_t cpg = C / G;
  const int es = in_dtype == RSM_F32 ? 4 : 2;
  // measured over BASELINE config 4's (C, G, D) grid: this form wins for groups of >= 8 channels, and for groups of 4
  // within one disparity chunk; narrower groups do so little arithmetic per output tile that the first kernel's
  // several-groups-per-staging-pass CTAs are faster
  if (cpg > GT_CT || cpg < 4 || (cpg < 8 && D > GT_DCH)) return RSM_ERR_UNSUPPORTED_CONFIG;
  if ((W * es) % 16 != 0) return RSM_ERR_UNSUPPORTED_CONFIG;
  GtGeom g;
  g.C = (int)C; g.G = (int)G; g.cpg = (int)cpg; g.H = (int)H; g.W = (int)W; g.D = (int)D;
  g.ndch = (int)ceil_div(D, GT_DCH);
  g.dch = (int)(ceil_div(ceil_div(D, g.ndch), 8) * 8);
  g.ntd = g.dch / GT_DT;
  g.gt = (int)(GT_CT / cpg < G ? GT_CT / cpg : G);
  g.gblocks = (int)ceil_div(G, g.gt);
  g.xtiles = (int)ceil_div(W, GT_TX);
  g.rw = GT_TX + g.dch;
  g.nsets = (32 * GT_CONS_WARPS) / (GT_NTX * g.ntd);
  if (g.nsets < 1) return RSM_ERR_UNSUPPORTED_CONFIG;
  if (g.nsets > g.gt) g.nsets = g.gt;
  g.pow2 = (cpg & (cpg - 1)) == 0;
  g.vec = D % 8 == 0 && aligned_to(out, 8 * (size_t)dtype_size(out_dtype));
  g.tiles = N * H * (int64_t)g.gblocks * g.ndch * g.xtiles;
  if (g.tiles > 2147483647LL) return RSM_ERR_UNSUPPORTED_CONFIG;
  g.l_bytes = g.tx_bytes = g.stage_bytes = g.nstage = 0;
  if (out_dtype != in_dtype && out_dtype != RSM_F32) return RSM_ERR_UNSUPPORTED_CONFIG;
  if (in_dtype == RSM_F32) return gt_launch<float, float>(left, right, out, N, g, st);
  if (in_dtype == RSM_F16) {
    if (out_dtype == RSM_F32) return gt_launch<__half, float>(left, right, out, N, g, st);
    return gt_launch<__half, __half>(left, right, out, N, g, st);
  }
  if (out_dtype == RSM_F32) return gt_launch<__nv_bfloat16, float>(left, right, out, N, g, st);
  return gt_launch<__nv_bfloat16, __nv_bfloat16>(left, right, out, N, g, st);
}

}  // namespace rsm
