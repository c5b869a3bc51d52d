// Correlation volumes (inner product / mean correlation / group-wise correlation), their adjoints,
// and the fused correlation -> soft-argmax/argmin kernel that never writes the volume.
//
// Forward tiling (SIMT, fp32 accumulate): a CTA owns one epipolar row segment of TX=64 pixels of one
// (n, y), a chunk of up to 64 disparities and a SLAB of up to 32 channels: either a slice of one big
// group (inner product: the CTA walks all slabs of the channel axis) or several whole small groups
// (group-wise: the CTA emits one output tile per group from a single staging pass).  The slab of
// the left segment and of the right window [x0-d_hi, x0+TX) is staged in shared memory as fp32
// once; each thread accumulates a 4(x) x 8(d) register tile from 4 LDS.128 per channel (one left
// quad + a 12-wide right window): 8 FMA per shared-memory word.  The lane order follows the
// output layout so that stores coalesce: x-fastest for (N,D,H,W), d-fastest with 256-bit stores
// (STG.E.ENL2.256) for the D-innermost (N,G,H,W,D).
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "rsm_common.cuh"

namespace rsm {

int launch_inner_tc(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                    int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st);
int launch_groupwise_tma(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                         int64_t D, int64_t G, int in_dtype, int out_dtype, cudaStream_t st);
int launch_inner_bwd_tc(const void* gout, const rsm_feat& left, const rsm_feat& right, void* gl, void* gr, int64_t N,
                        int64_t C, int64_t H, int64_t W, int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st,
                        unsigned long long* prof);
int launch_inner_regress_rows(const rsm_feat& left, const rsm_feat& right, int64_t N, int64_t C, int64_t H, int64_t W,
                              int64_t D, int mean, int in_dtype, const rsm_regress_out& out, cudaStream_t st,
                              unsigned long long* prof);

constexpr int TX = 64;        // pixels per CTA tile
constexpr int XT = 4;         // pixels per thread
constexpr int DT = 8;         // disparities per thread
constexpr int NTX = TX / XT;  // 16 threads along x
constexpr int MAX_NTD = 8;    // <= 64 disparities per CTA chunk (volume kernels)
constexpr int CKMAX = 32;     // channels per shared-memory slab

enum Layout { LAYOUT_NDHW = 0, LAYOUT_NGHWD = 1 };

struct CorrGeom {
  int C, H, W, D, G, cpg;
  int ntd;      // threads along d  (blockDim.x = NTX * ntd)
  int dchp;     // disparities covered by a CTA chunk = DT * ntd
  int xtiles;   // ceil(W / TX)
  int mean;     // divide by cpg
  int gpb;      // groups per CTA (small groups share one staging pass)
  int gblocks;  // ceil(G / gpb)
  int pow2;     // cpg is a power of two: the mean is an exact multiply
  int pairs;    // 16-bit features can be staged as aligned 4-byte pairs with cp.async
};

template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<float>(uint32_t w) { return make_float2(0.f, 0.f); }   // unused
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// ---- stage `nch` channels [c0, c0+nch) of the left segment and right window as fp32
template <typename Tin>
__device__ __forceinline__ void stage_slab(const FeatView& L, const FeatView& R, int64_t n, int y, int c0, int nch,
                                           int x0, int rbase, int rw, float* sL, float* sR, int W,
                                           uint32_t* sRaw = nullptr) {
  const Tin* __restrict__ pl = reinterpret_cast<const Tin*>(L.data) + n * L.sn + (int64_t)y * L.sh + (int64_t)c0 * L.sc;
  const Tin* __restrict__ pr = reinterpret_cast<const Tin*>(R.data) + n * R.sn + (int64_t)y * R.sh + (int64_t)c0 * R.sc;
  const int span = TX + rw;   // per channel: TX left values then rw right values
  if constexpr (sizeof(Tin) == 2) {
    if (sRaw) {
      // 16-bit features, pair-aligned: raw 4-byte pairs go global -> shared with cp.async (every load of the
      // slab in flight at once, zero-fill outside the image), then one conversion pass widens them to fp32
      const int hspan = span / 2;
      for (int e = threadIdx.x; e < hspan; e += blockDim.x) {
        const bool left = e < TX / 2;
        const int x = left ? x0 + 2 * e : rbase + 2 * (e - TX / 2);
        const bool valid = x >= 0 && x + 1 < W;
        const Tin* src = valid ? (left ? pl + x : pr + x) : reinterpret_cast<const Tin*>(L.data);
        const int64_t step = valid ? (left ? L.sc : R.sc) : 0;
        const int nbytes = valid ? 4 : 0;
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(sRaw + e);
        for (int c = 0; c < nch; ++c)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sdst + (uint32_t)(c * hspan * 4)),
                       "l"(src + c * step), "r"(nbytes)
                       : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      for (int c = 0; c < nch; ++c)
        for (int j = threadIdx.x; j < hspan; j += blockDim.x) {
          const float2 f = unpack2<Tin>(sRaw[c * hspan + j]);
          float* dst = j < TX / 2 ? sL + c * TX + 2 * j : sR + c * rw + 2 * (j - TX / 2);
          *reinterpret_cast<float2*>(dst) = f;
        }
      return;
    }
  }
  constexpr int U = 8;        // independent loads in flight per thread
  for (int e = threadIdx.x; e < span; e += blockDim.x) {
    const bool left = e < TX;
    const int x = left ? x0 + e : rbase + (e - TX);
    const bool valid = x >= 0 && x < W;
    const Tin* __restrict__ p = left ? pl + (int64_t)x * L.sw : pr + (int64_t)x * R.sw;
    const int64_t cs = left ? L.sc : R.sc;
    float* dst = left ? sL + e : sR + (e - TX);
    const int pitch = left ? TX : rw;
    if constexpr (sizeof(Tin) == 4) {
      // fp32 features: asynchronous global->shared copies (LDGSTS), all of the thread's loads in
      // flight at once, zero-fill (src-size 0) outside the image; completed by cp_async_wait_all()
      const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst);
      const Tin* src = valid ? p : reinterpret_cast<const Tin*>(L.data);
      const int64_t step = valid ? cs : 0;
      const int nbytes = valid ? 4 : 0;
      for (int c = 0; c < nch; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sdst + (uint32_t)(c * pitch * 4)),
                     "l"(src + c * step), "r"(nbytes)
                     : "memory");
      continue;
    }
    for (int c = 0; c < nch; c += U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (valid && c + u < nch) ? to_f(__ldg(p + (c + u) * cs)) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c + u < nch) dst[(c + u) * pitch] = v[u];
    }
  }
  if constexpr (sizeof(Tin) == 4) asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- accumulate the 4x8 register tile over staged channels [cofs, cofs+nch)
__device__ __forceinline__ void tile_fma(const float* sL, const float* sR, int rw, int tx, int wstart, int cofs,
                                         int nch, float (&acc)[XT][DT]) {
  const float* pl = sL + cofs * TX + XT * tx;
  const float* pw = sR + cofs * rw + wstart;
#pragma unroll 4
  for (int c = 0; c < nch; ++c) {
    const float4 l4 = *reinterpret_cast<const float4*>(pl + c * TX);
    const float4* wp = reinterpret_cast<const float4*>(pw + c * rw);
    const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const float l[XT] = {l4.x, l4.y, l4.z, l4.w};
    const float w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
      for (int j = 0; j < DT; ++j) acc[i][j] = fmaf(l[i], w[8 + i - j], acc[i][j]);
  }
}

// ---- vector stores of consecutive outputs (streaming: the volume is written once, read later)
__device__ __forceinline__ void store4(float* p, const float* v) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void store4(__half* p, const float* v) {
  union { uint2 u; __half2 h[2]; } t;
  t.h[0] = __floats2half2_rn(v[0], v[1]); t.h[1] = __floats2half2_rn(v[2], v[3]);
  __stcs(reinterpret_cast<uint2*>(p), t.u);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float* v) {
  union { uint2 u; __nv_bfloat162 h[2]; } t;
  t.h[0] = __floats2bfloat162_rn(v[0], v[1]); t.h[1] = __floats2bfloat162_rn(v[2], v[3]);
  __stcs(reinterpret_cast<uint2*>(p), t.u);
}
// 8 consecutive outputs: one 256-bit store for fp32 (32-byte aligned), one 128-bit store for 16-bit
__device__ __forceinline__ void store8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void store8(__half* p, const float* v) {
  union { uint4 u; __half2 h[4]; } t;
#pragma unroll
  for (int k = 0; k < 4; ++k) t.h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
  __stcs(reinterpret_cast<uint4*>(p), t.u);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
  union { uint4 u; __nv_bfloat162 h[4]; } t;
#pragma unroll
  for (int k = 0; k < 4; ++k) t.h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  __stcs(reinterpret_cast<uint4*>(p), t.u);
}

// ---- scale, zero the x < d triangle and store one finished 4x8 tile of group `grp`
template <typename Tout, int LAYOUT>
__device__ __forceinline__ void store_tile(float (&acc)[XT][DT], Tout* __restrict__ out, const CorrGeom& g, int64_t n,
                                           int grp, int y, int xb, int db, bool vec) {
  if (g.mean) {
    if (g.pow2) {
      const float inv = 1.f / (float)g.cpg;   // exact: identical to the division
#pragma unroll
      for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[i][j] *= inv;
    } else {
      const float cnt = (float)g.cpg;
#pragma unroll
      for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[i][j] = acc[i][j] / cnt;
    }
  }
  if (xb < db + DT - 1) {                      // only tiles touching the x < d triangle pay for the masking
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
      for (int j = 0; j < DT; ++j)
        if (xb + i < db + j) acc[i][j] = 0.f;   // the reference leaves zeros where x < d
  }

  if constexpr (LAYOUT == LAYOUT_NDHW) {
    if (xb >= g.W) return;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      const int d = db + j;
      if (d >= g.D) break;
      Tout* p = out + (((int64_t)n * g.D + d) * g.H + y) * g.W + xb;
      if (vec) {   // W % 4 == 0: xb + 3 < W and the address is aligned
        const float v[4] = {acc[0][j], acc[1][j], acc[2][j], acc[3][j]};
        store4(p, v);
      } else {
#pragma unroll
        for (int i = 0; i < XT; ++i)
          if (xb + i < g.W) p[i] = from_f<Tout>(acc[i][j]);
      }
    }
  } else {
    if (db >= g.D) return;
#pragma unroll
    for (int i = 0; i < XT; ++i) {
      const int x = xb + i;
      if (x >= g.W) break;
      Tout* p = out + ((((int64_t)n * g.G + grp) * g.H + y) * g.W + x) * g.D + db;
      if (vec) {   // D % 8 == 0: db + 7 < D and the address is aligned
        store8(p, &acc[i][0]);
      } else {
#pragma unroll
        for (int j = 0; j < DT; ++j)
          if (db + j < g.D) p[j] = from_f<Tout>(acc[i][j]);
      }
    }
  }
}

// ======================================================================= volume forward
template <typename Tin, typename Tout, int LAYOUT>
__global__ void __launch_bounds__(NTX * MAX_NTD)
corr_fwd_kernel(FeatView L, FeatView R, Tout* __restrict__ out, CorrGeom g, int vec) {
  extern __shared__ __align__(16) float smem[];
  const int rw = TX + g.dchp;
  float* sL = smem;
  float* sR = smem + CKMAX * TX;
  uint32_t* sRaw = g.pairs ? reinterpret_cast<uint32_t*>(sR + CKMAX * rw) : nullptr;

  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % g.xtiles); bid /= g.xtiles;
  const int gb = (int)(bid % g.gblocks); bid /= g.gblocks;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * TX;
  const int dc0 = blockIdx.y * g.dchp;
  int tx, td;
  if constexpr (LAYOUT == LAYOUT_NDHW) { tx = threadIdx.x % NTX; td = threadIdx.x / NTX; }
  else { td = threadIdx.x % g.ntd; tx = threadIdx.x / g.ntd; }
  const int rbase = x0 - dc0 - g.dchp;
  const int wstart = g.dchp + XT * tx - DT * td - DT;
  const int xb = x0 + XT * tx, db = dc0 + DT * td;

  float acc[XT][DT];
  const int g0 = gb * g.gpb, g1 = min(g.G, g0 + g.gpb);
  if (g.cpg >= CKMAX) {
    // one (big) group per CTA: walk its channels slab by slab into the same accumulators
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
      for (int j = 0; j < DT; ++j) acc[i][j] = 0.f;
    const int cbeg = g0 * g.cpg, cend = cbeg + g.cpg;
    for (int c0 = cbeg; c0 < cend; c0 += CKMAX) {
      const int nch = min(CKMAX, cend - c0);
      __syncthreads();
      stage_slab<Tin>(L, R, n, y, c0, nch, x0, rbase, rw, sL, sR, g.W, sRaw);
      __syncthreads();
      tile_fma(sL, sR, rw, tx, wstart, 0, nch, acc);
    }
    store_tile<Tout, LAYOUT>(acc, out, g, n, g0, y, xb, db, vec);
  } else {
    // several small groups per CTA: one staging pass, one output tile per group
    stage_slab<Tin>(L, R, n, y, g0 * g.cpg, (g1 - g0) * g.cpg, x0, rbase, rw, sL, sR, g.W, sRaw);
    __syncthreads();
    for (int grp = g0; grp < g1; ++grp) {
#pragma unroll
      for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[i][j] = 0.f;
      tile_fma(sL, sR, rw, tx, wstart, (grp - g0) * g.cpg, g.cpg, acc);
      store_tile<Tout, LAYOUT>(acc, out, g, n, grp, y, xb, db, vec);
    }
  }
}

// ============================================ inner product, fp32 features: 8(x) x 16(d) register tiles
// The 4x8 tile above needs 4 LDS.128 per 32 FMA: with four warps issuing FMAs the shared-memory pipe (one
// 128-byte wavefront per clock) saturates at about half the FMA rate.  For the (N,D,H,W) inner product on
// fp32 features -- the reference's training dtype, too wide for one tensor-core pass -- this kernel doubles
// the arithmetic per byte read from shared memory: CTA = 128 pixels x <= 64 disparities of one (n, y),
// thread tile 8(x) x 16(d) = 128 accumulators fed by 2 + 6 LDS.128 per channel (16 FMA per LDS.128).
// Threads of a warp read 32-byte segments 32 bytes apart; the 16-byte chunks of every staged row are
// XOR-swizzled (chunk ^= (chunk >> 3) & 1) so that the two halves of a quarter-warp land in different banks.
// Staging: 16-byte LDGSTS with zero-fill (rows and windows start on 16-byte boundaries: fast path only).
// The disparity extent of a thread tile is a template parameter (8, 12 or 16) chosen so that 16 * ntd threads
// fill whole warps for the usual D (48 = 4 x 12, 64 = 4 x 16, 24 = 2 x 12, 16 = 2 x 8).
constexpr int BG_TX = 128, BG_XT = 8, BG_NTX = BG_TX / BG_XT, BG_CK = 32, BG_MAXNTD = 4;

__device__ __forceinline__ int bg_swz(int chunk) { return chunk ^ ((chunk >> 3) & 1); }

struct RegressOutPtrs {   // FUSED: the tile goes to the soft-argmax / arg-extrema scan instead of global memory
  float* soft;
  int64_t* amin;
  int64_t* amax;
  float* lse;
};

template <typename Tout, int BG_DT, bool FUSED>
__global__ void __launch_bounds__(BG_NTX * BG_MAXNTD)
inner_fwd_big_kernel(FeatView L, FeatView R, Tout* __restrict__ out, CorrGeom g, int dchp, int xtiles, int vec8,
                     RegressOutPtrs ro) {
  extern __shared__ __align__(16) float smem[];
  const int rw = BG_TX + dchp;                     // right window width (multiple of 16 floats)
  float* sL = smem;                                // [BG_CK][BG_TX]
  float* sR = smem + BG_CK * BG_TX;                // [BG_CK][rw]
  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % xtiles); bid /= xtiles;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * BG_TX, dc0 = blockIdx.y * dchp;
  const int tx = threadIdx.x % BG_NTX, td = threadIdx.x / BG_NTX;
  const int rbase = x0 - dc0 - dchp;               // first pixel of the right window
  const int wchunk = (dchp + BG_XT * tx - BG_DT * td - BG_DT) >> 2;   // window chunk of this thread: w[k], k = DT + i - j
  const float* __restrict__ pl = reinterpret_cast<const float*>(L.data) + n * L.sn + (int64_t)y * L.sh;
  const float* __restrict__ pr = reinterpret_cast<const float*>(R.data) + n * R.sn + (int64_t)y * R.sh;

  float acc[BG_XT][BG_DT];
#pragma unroll
  for (int i = 0; i < BG_XT; ++i)
#pragma unroll
    for (int j = 0; j < BG_DT; ++j) acc[i][j] = 0.f;

  const int lch = BG_TX / 4, rch = rw / 4;          // chunks per row
  for (int c0 = 0; c0 < g.C; c0 += BG_CK) {
    const int nch = min(BG_CK, g.C - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < lch + rch; e += blockDim.x) {
      const bool left = e < lch;
      const int ch = left ? e : e - lch;
      const int x = left ? x0 + 4 * ch : rbase + 4 * ch;
      const bool valid = x >= 0 && x < g.W;          // W % 4 == 0 and x % 4 == 0: whole chunks
      const float* src = valid ? (left ? pl : pr) + (int64_t)c0 * (left ? L.sc : R.sc) + x : pl;
      const int64_t step = valid ? (left ? L.sc : R.sc) : 0;
      const uint32_t sdst = (uint32_t)__cvta_generic_to_shared((left ? sL : sR) + 4 * bg_swz(ch));
      const uint32_t pitch = (uint32_t)(left ? BG_TX : rw) * 4u;
      const int nbytes = valid ? 16 : 0;
      for (int c = 0; c < nch; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + c * pitch), "l"(src + c * step), "r"(nbytes)
                     : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const float* ql = sL;
    const float* qr = sR;
#pragma unroll 2
    for (int c = 0; c < nch; ++c, ql += BG_TX, qr += rw) {
      const float4 l0 = *reinterpret_cast<const float4*>(ql + 4 * bg_swz(2 * tx));
      const float4 l1 = *reinterpret_cast<const float4*>(ql + 4 * bg_swz(2 * tx + 1));
      constexpr int NW = (BG_DT + BG_XT) / 4;     // chunks covering k = 1 .. DT + 7
      float w[4 * NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const float4 t = *reinterpret_cast<const float4*>(qr + 4 * bg_swz(wchunk + k));
        w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
      }
      const float l[BG_XT] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int i = 0; i < BG_XT; ++i)
#pragma unroll
        for (int j = 0; j < BG_DT; ++j) acc[i][j] = fmaf(l[i], w[BG_DT + i - j], acc[i][j]);
    }
  }
  const int xb = x0 + BG_XT * tx, db = dc0 + BG_DT * td;
  const float inv = 1.f / (float)g.cpg, cnt = (float)g.cpg;
  if constexpr (FUSED) {
    // ---- the whole disparity range is in this CTA (D <= dchp): park the scaled tile over the dead operand slabs,
    // sV[x][d] with an odd pitch, then one thread per pixel walks its D values in ascending order: chunked online
    // softmax (one rescale exp per 8 values), strict compares so the first index wins ties, NaNs win through a
    // flag -- torch.argmin / argmax semantics (same scan as inner_regress_fwd_kernel)
    const int dp = dchp + 1;
    float* sV = smem;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BG_XT; ++i)
#pragma unroll
      for (int j = 0; j < BG_DT; ++j) {
        float a = acc[i][j];
        if (g.mean) a = g.pow2 ? a * inv : a / cnt;
        sV[(BG_XT * tx + i) * dp + BG_DT * td + j] = xb + i < db + j ? 0.f : a;
      }
    __syncthreads();
    for (int xx = threadIdx.x; xx < BG_TX; xx += blockDim.x) {
      const int x = x0 + xx;
      if (x >= g.W) continue;
      const float* col = sV + xx * dp;
      ScanState sc;
      int d0 = 0;
      for (; d0 + 8 <= g.D; d0 += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = col[d0 + k];
        sc.chunk8(v, d0);
      }
      if (d0 < g.D) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = d0 + k < g.D ? col[d0 + k] : 0.f;
        sc.tail(v, d0, g.D - d0);
      }
      sc.finish();
      const int64_t o = ((int64_t)n * g.H + y) * g.W + x;
      if (ro.soft) ro.soft[o] = sc.ws / sc.s;
      if (ro.lse) ro.lse[o] = sc.m + __logf(sc.s);
      if (ro.amin) ro.amin[o] = sc.mini;
      if (ro.amax) ro.amax[o] = sc.maxi;
    }
    return;
  }
  // ---- scale, zero the x < d triangle, store: 8 consecutive pixels per disparity
  if (xb >= g.W) return;
#pragma unroll
  for (int j = 0; j < BG_DT; ++j) {
    const int d = db + j;
    if (d >= g.D) break;
    float v[BG_XT];
#pragma unroll
    for (int i = 0; i < BG_XT; ++i) {
      float a = acc[i][j];
      if (g.mean) a = g.pow2 ? a * inv : a / cnt;
      v[i] = xb + i < d ? 0.f : a;
    }
    Tout* p = out + (((int64_t)n * g.D + d) * g.H + y) * g.W + xb;
    if (vec8) store8(p, v);
    else {                                             // W % 4 == 0 and xb % 8 == 0: whole quads
      store4(p, v);
      if (xb + 4 < g.W) store4(p + 4, v + 4);
    }
  }
}

// ======================================================================= volume adjoint
// One thread per (n,c,y,x); atomic-free gather over d (SURVEY.md 8a backward contracts):
//   gL[c,x]  = s * sum_{d<=min(x,D-1)}       gV[g(c),d,x]    * R[c,x-d]
//   gR[c,x'] = s * sum_{d<=min(D-1,W-1-x')}  gV[g(c),d,x'+d] * L[c,x'+d]
template <typename Tin, typename Tout, int LAYOUT>
__global__ void __launch_bounds__(256)
corr_bwd_kernel(const Tout* __restrict__ gout, FeatView L, FeatView R, Tin* __restrict__ gl,
                Tin* __restrict__ gr, int64_t total, CorrGeom g) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % g.W);
  const int y = (int)((i / g.W) % g.H);
  const int c = (int)((i / ((int64_t)g.W * g.H)) % g.C);
  const int64_t n = i / ((int64_t)g.W * g.H * g.C);
  const int grp = c / g.cpg;
  const Tin* __restrict__ pl = reinterpret_cast<const Tin*>(L.data) + n * L.sn + (int64_t)c * L.sc + (int64_t)y * L.sh;
  const Tin* __restrict__ pr = reinterpret_cast<const Tin*>(R.data) + n * R.sn + (int64_t)c * R.sc + (int64_t)y * R.sh;
  // address of gV[d, x]:  base + d * sd + x * sx
  int64_t base, sd, sx;
  if constexpr (LAYOUT == LAYOUT_NDHW) {
    base = ((int64_t)n * g.D * g.H + y) * g.W; sd = (int64_t)g.H * g.W; sx = 1;
  } else {
    base = (((int64_t)n * g.G + grp) * g.H + y) * (int64_t)g.W * g.D; sd = 1; sx = g.D;
  }
  const float cnt = g.mean ? (float)g.cpg : 1.f;
  if (gl) {
    float s = 0.f;
    const int dl = min(x, g.D - 1);
    for (int d = 0; d <= dl; ++d)
      s = fmaf(to_f(__ldg(gout + base + d * sd + x * sx)), to_f(__ldg(pr + (int64_t)(x - d) * R.sw)), s);
    gl[i] = from_f<Tin>(s / cnt);
  }
  if (gr) {
    float s = 0.f;
    const int dr = min(g.D - 1, g.W - 1 - x);
    for (int d = 0; d <= dr; ++d)
      s = fmaf(to_f(__ldg(gout + base + d * sd + (x + d) * sx)), to_f(__ldg(pl + (int64_t)(x + d) * L.sw)), s);
    gr[i] = from_f<Tin>(s / cnt);
  }
}

// ============================================================ fused correlation -> regression
struct Best {   // (value, index) with torch arg-extremum ordering
  float v; int i;
};
template <bool MIN> __device__ __forceinline__ Best better(Best a, Best b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) {
    if (an && bn) return a.i < b.i ? a : b;
    return an ? a : b;
  }
  if (MIN ? (a.v < b.v) : (a.v > b.v)) return a;
  if (MIN ? (a.v > b.v) : (a.v < b.v)) return b;
  return a.i < b.i ? a : b;
}
template <bool MIN> __device__ __forceinline__ Best warp_best(Best a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better<MIN>(a, b);
  }
  return a;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The CTA covers ALL disparities of a TX-pixel row segment (blockDim = NTX * ceil(D/8) <= 1024),
// parks the fp32 tile in shared memory and reduces each pixel's D values with one warp:
// max -> (sum exp, sum d*exp) -> argmin/argmax, so the (N,D,H,W) volume never reaches HBM.
template <typename Tin, int MAXT>
__global__ void __launch_bounds__(MAXT)
inner_regress_fwd_kernel(FeatView L, FeatView R, float* __restrict__ soft, int64_t* __restrict__ amin,
                         int64_t* __restrict__ amax, float* __restrict__ lse, CorrGeom g) {
  extern __shared__ __align__(16) float smem[];
  const int rw = TX + g.dchp;
  const int dp = g.dchp + 1;                 // odd pitch of the parked tile sV[x][d]
  float* sL = smem;
  float* sR = sL + CKMAX * TX;
  float* sV = sR + CKMAX * rw;

  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % g.xtiles); bid /= g.xtiles;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * TX;
  const int tx = threadIdx.x % NTX, td = threadIdx.x / NTX;
  const int rbase = x0 - g.dchp;
  const int wstart = g.dchp + XT * tx - DT * td - DT;

  float acc[XT][DT];
#pragma unroll
  for (int i = 0; i < XT; ++i)
#pragma unroll
    for (int j = 0; j < DT; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < g.C; c0 += CKMAX) {
    const int nch = min(CKMAX, g.C - c0);
    __syncthreads();
    stage_slab<Tin>(L, R, n, y, c0, nch, x0, rbase, rw, sL, sR, g.W);
    __syncthreads();
    tile_fma(sL, sR, rw, tx, wstart, 0, nch, acc);
  }
  const float cnt = g.mean ? (float)g.C : 1.f, inv = 1.f / cnt;
  const bool pow2 = !g.mean || g.pow2;            // exact reciprocal: identical to the division
  const int xb = XT * tx, db = DT * td;
#pragma unroll
  for (int i = 0; i < XT; ++i)
#pragma unroll
    for (int j = 0; j < DT; ++j)
      sV[(xb + i) * dp + db + j] = (x0 + xb + i >= db + j) ? (pow2 ? acc[i][j] * inv : acc[i][j] / cnt) : 0.f;
  __syncthreads();

  // One thread per pixel column, walking its D values in ascending order from shared memory (odd pitch:
  // conflict-free): chunked online softmax (one rescale exp per 8 values), strict compares so the
  // first index wins ties, NaNs win through a flag -- torch.argmin/argmax semantics.
  for (int xx = threadIdx.x; xx < TX; xx += blockDim.x) {
    const int x = x0 + xx;
    if (x >= g.W) continue;
    const float* col = sV + xx * dp;
    ScanState sc;
    int d0 = 0;
    for (; d0 + 8 <= g.D; d0 += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = col[d0 + k];
      sc.chunk8(v, d0);
    }
    if (d0 < g.D) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = d0 + k < g.D ? col[d0 + k] : 0.f;
      sc.tail(v, d0, g.D - d0);
    }
    sc.finish();
    const int64_t o = ((int64_t)n * g.H + y) * g.W + x;
    if (soft) soft[o] = sc.ws / sc.s;
    if (lse) lse[o] = sc.m + __logf(sc.s);
    if (amin) amin[o] = sc.mini;
    if (amax) amax[o] = sc.maxi;
  }
}

static bool grid_ok(int64_t blocks) { return blocks >= 0 && blocks <= 2147483647LL; }

static int make_geom(int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G, int mean, bool all_d,
                     CorrGeom& g) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0 || G <= 0) return RSM_ERR_INVALID_SHAPE;
  if (C % G != 0) return RSM_ERR_INVALID_SHAPE;
  if (C > (1 << 24) || H > (1 << 24) || W > (1 << 24) || D > (1 << 24)) return RSM_ERR_INVALID_SHAPE;
  g.C = (int)C; g.H = (int)H; g.W = (int)W; g.D = (int)D; g.G = (int)G; g.cpg = (int)(C / G);
  const int need = (int)ceil_div(D > 0 ? D : 1, DT);
  // the fused kernel reduces with full-warp shuffles: keep blockDim = 16 * ntd a multiple of 32
  g.ntd = all_d ? ((need + 1) & ~1) : (need < MAX_NTD ? need : MAX_NTD);
  g.dchp = g.ntd * DT;
  g.xtiles = (int)ceil_div(W > 0 ? W : 1, TX);
  g.mean = mean;
  g.gpb = (g.cpg >= CKMAX || g.cpg == 0) ? 1 : (CKMAX / g.cpg < g.G ? CKMAX / g.cpg : g.G);
  g.gblocks = (int)ceil_div(g.G, g.gpb);
  g.pow2 = g.cpg > 0 && (g.cpg & (g.cpg - 1)) == 0;
  g.pairs = 0;
  return RSM_OK;
}

template <typename Tin, typename Tout, int LAYOUT>
static int launch_fwd(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, const CorrGeom& g_in,
                      cudaStream_t st, const char* where) {
  CorrGeom g = g_in;
  const int64_t bx = N * g.gblocks * g.H * g.xtiles;
  const int64_t by = ceil_div(g.D, g.dchp);
  if (!grid_ok(bx) || by > 65535) return RSM_ERR_INVALID_SHAPE;
  auto pair_ok = [&](const rsm_feat& f) {
    return f.stride_w == 1 && f.stride_n % 2 == 0 && f.stride_c % 2 == 0 && f.stride_h % 2 == 0 && aligned_to(f.data, 4);
  };
  g.pairs = sizeof(Tin) == 2 && g.W % 2 == 0 && pair_ok(left) && pair_ok(right);
  const size_t smem = (size_t)(CKMAX * TX + CKMAX * (TX + g.dchp)) * sizeof(float) +
                      (g.pairs ? (size_t)CKMAX * (2 * TX + g.dchp) * 2 : 0);
  // vector stores need the run length to divide evenly and the base pointer aligned to the vector
  const int vec = (LAYOUT == LAYOUT_NDHW) ? (g.W % 4 == 0 && aligned_to(out, 4 * sizeof(Tout)))
                                          : (g.D % 8 == 0 && aligned_to(out, 8 * sizeof(Tout)));
  corr_fwd_kernel<Tin, Tout, LAYOUT><<<dim3((unsigned)bx, (unsigned)by), NTX * g.ntd, smem, st>>>(
      view_of(left), view_of(right), (Tout*)out, g, vec);
  return finish_launch(where);
}

}  // namespace rsm
#include "rsm_corr_bwd.cuh"
#include "rsm_groupwise_bwd.cuh"
namespace rsm {

template <typename Tin, typename Tout, int LAYOUT>
static int launch_bwd(const void* gout, const rsm_feat& left, const rsm_feat& right, void* gl, void* gr,
                      int64_t N, const CorrGeom& g, cudaStream_t st, const char* where) {
  // tiled kernels: channel blocks of <= 32 channels that never straddle a group
  int cbs = 0;
  for (int c = BW_CB; c >= 4; c -= 4)
    if (g.cpg % c == 0) { cbs = c; break; }
  // D-innermost volume, 1..16 channels per group: one CTA per (n, group, y) row streams the gradient row once through
  // 16-disparity slabs and produces both gradients from it (rsm_groupwise_bwd.cuh)
  if constexpr (LAYOUT == LAYOUT_NGHWD) {
    int rc = RSM_OK;
    if (launch_groupwise_bwd_slab<Tin, Tout>(gout, left, right, gl, gr, N, g, st, where, rc)) return rc;
  }
  // inner product: 8(x) x 4(c) register tiles when every row and window starts on a 16-byte boundary (fp32: W % 4,
  // 16-bit: W % 8) and all three tensors share one dtype
  if constexpr (LAYOUT == LAYOUT_NDHW && std::is_same<Tin, Tout>::value) {
    constexpr int EPV = 16 / (int)sizeof(Tin);
    auto v16 = [&](const rsm_feat& f) {
      return f.stride_w == 1 && f.stride_n % EPV == 0 && f.stride_c % EPV == 0 && f.stride_h % EPV == 0 && aligned_to(f.data, 16);
    };
    // (16-bit tensors are widened synchronously while staging, a few loads in flight per thread)
    if (g.G == 1 && g.C >= 16 && g.D > 0 && g.W % EPV == 0 && v16(left) && v16(right) && aligned_to(gout, 16) &&
        (!gl || aligned_to(gl, 16)) && (!gr || aligned_to(gr, 16))) {
      const int xtiles = (int)ceil_div(g.W, BB_TX), cblocks = (int)ceil_div(g.C, BB_CB);
      const int64_t bx = N * g.H * (int64_t)cblocks * xtiles;
      if (grid_ok(bx)) {
        // disparities per pass: 48 unless 64 leaves fewer dead rows in the last pass
        const int64_t waste48 = ceil_div(g.D, 48) * 48 - g.D, waste64 = ceil_div(g.D, 64) * 64 - g.D;
        const bool d48 = waste48 <= waste64;
        const size_t smem = bb_smem_bytes(d48 ? 48 : 64);
        auto launch = [&](auto k, void* dst) -> int {
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          k<<<(unsigned)bx, BB_THREADS, smem, st>>>((const Tin*)gout, view_of(left), view_of(right), (Tin*)dst, g, xtiles, cblocks);
          return finish_launch(where);
        };
        if (gl) {
          if (int rc = d48 ? launch(inner_bwd_big_kernel<Tin, SIDE_LEFT, 48>, gl) : launch(inner_bwd_big_kernel<Tin, SIDE_LEFT, 64>, gl)) return rc;
        }
        if (gr) {
          if (int rc = d48 ? launch(inner_bwd_big_kernel<Tin, SIDE_RIGHT, 48>, gr) : launch(inner_bwd_big_kernel<Tin, SIDE_RIGHT, 64>, gr)) return rc;
        }
        return RSM_OK;
      }
    }
  }
  if (cbs > 0 && g.D > 0) {
    const int cblocks = g.C / cbs;
    const int64_t bx = N * g.H * (int64_t)cblocks * g.xtiles;
    if (!grid_ok(bx)) return RSM_ERR_INVALID_SHAPE;
    if (gl) {
      const size_t smem = (size_t)(BW_DCH * (BW_TX + 4) + cbs * (BW_TX + BW_DCH)) * sizeof(float);
      corr_bwd_tiled_kernel<Tin, Tout, LAYOUT, SIDE_LEFT><<<(unsigned)bx, 128, smem, st>>>(
          (const Tout*)gout, view_of(left), view_of(right), (Tin*)gl, g, cblocks, cbs);
      if (int rc = finish_launch(where)) return rc;
    }
    if (gr) {
      const size_t smem = (size_t)(BW_DCH * (BW_TX + BW_DCH + 4) + cbs * (BW_TX + BW_DCH)) * sizeof(float);
      auto k = corr_bwd_tiled_kernel<Tin, Tout, LAYOUT, SIDE_RIGHT>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<(unsigned)bx, 128, smem, st>>>((const Tout*)gout, view_of(left), view_of(right), (Tin*)gr, g, cblocks, cbs);
      if (int rc = finish_launch(where)) return rc;
    }
    return RSM_OK;
  }
  const int64_t total = N * g.C * g.H * g.W;
  if (!grid_ok(ceil_div(total, 256))) return RSM_ERR_INVALID_SHAPE;
  corr_bwd_kernel<Tin, Tout, LAYOUT><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
      (const Tout*)gout, view_of(left), view_of(right), (Tin*)gl, (Tin*)gr, total, g);
  return finish_launch(where);
}

// dispatch on (in_dtype, out_dtype): outputs are either the input dtype or fp32
#define RSM_DISPATCH_IO(in_dt, out_dt, Tin, Tout, ...)                                   \
  [&]() -> int {                                                                         \
    if (out_dt != in_dt && out_dt != RSM_F32) return (int)RSM_ERR_UNSUPPORTED_DTYPE;     \
    return RSM_DISPATCH_DTYPE(in_dt, Tin, [&]() -> int {                                 \
      if (out_dt == RSM_F32) { using Tout = float; return __VA_ARGS__(); }               \
      using Tout = Tin;                                                                  \
      return __VA_ARGS__();                                                              \
    });                                                                                  \
  }()

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

extern "C" int rsm_inner_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                             int64_t W, int64_t D, int reduce, int in_dtype, int out_dtype, int device,
                             void* stream) {
  CorrGeom g;
  if (int rc = make_geom(N, C, H, W, D, 1, reduce == RSM_REDUCE_MEAN, false, g)) return rc;
  if (N * H * W * D == 0) return RSM_OK;
  if (!out || (C > 0 && (!left.data || !right.data))) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(in_dtype)
  if (!valid_dtype(out_dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (!aligned_to(out, dtype_size(out_dtype))) return RSM_ERR_MISALIGNED;
  // 16-bit features: banded per-row GEMM on the tcgen05 tensor cores (rsm_corr_tc.cu); fp32 features and
  // shapes it does not cover (C % 16 != 0) use the SIMT kernel.
  {
    const int rc = launch_inner_tc(left, right, out, N, C, H, W, D, reduce == RSM_REDUCE_MEAN, in_dtype, out_dtype, st);
    if (rc != RSM_ERR_UNSUPPORTED_CONFIG) return rc;
  }
  // fp32 features: 8x16 register tiles when rows and windows start on 16-byte boundaries
  {
    auto v4 = [&](const rsm_feat& f) {
      return f.stride_w == 1 && f.stride_n % 4 == 0 && f.stride_c % 4 == 0 && f.stride_h % 4 == 0 && aligned_to(f.data, 16);
    };
    if (in_dtype == RSM_F32 && out_dtype == RSM_F32 && C >= 16 && D >= 16 && W % 4 == 0 && v4(left) && v4(right) &&
        aligned_to(out, 16)) {
      // (DT, ntd) with an even ntd (whole warps): the smallest chunk DT * ntd covering min(D, 64)
      static const int kDT[5] = {8, 12, 8, 12, 16}, kNTD[5] = {2, 2, 4, 4, 4};   // chunks 16, 24, 32, 48, 64
      int pick = 4;
      for (int k = 0; k < 5; ++k)
        if (kDT[k] * kNTD[k] >= (D < 64 ? D : 64)) { pick = k; break; }
      // several chunks: prefer the split with the least padding (e.g. D = 96 -> 2 x 48, not 64 + 32 of 64)
      if (D > 64) {
        int64_t best = -1;
        for (int k = 2; k < 5; ++k) {
          const int64_t ch = kDT[k] * kNTD[k], padded = ceil_div(D, ch) * ch;
          if (best < 0 || padded <= best) { best = padded; pick = k; }   // ties: the wider tile
        }
      }
      const int dt = kDT[pick], dchp = dt * kNTD[pick];
      const int xtiles = (int)ceil_div(W, BG_TX);
      const int64_t bx = N * H * xtiles, by = ceil_div(D, dchp);
      if (grid_ok(bx) && by <= 65535) {
        const size_t smem = (size_t)BG_CK * (2 * BG_TX + dchp) * sizeof(float);
        const int vec8 = W % 8 == 0 && aligned_to(out, 32);
        const dim3 grid((unsigned)bx, (unsigned)by);
        const unsigned nt = BG_NTX * kNTD[pick];
        const RegressOutPtrs none{nullptr, nullptr, nullptr, nullptr};
        if (dt == 8) inner_fwd_big_kernel<float, 8, false><<<grid, nt, smem, st>>>(view_of(left), view_of(right), (float*)out, g, dchp, xtiles, vec8, none);
        else if (dt == 12) inner_fwd_big_kernel<float, 12, false><<<grid, nt, smem, st>>>(view_of(left), view_of(right), (float*)out, g, dchp, xtiles, vec8, none);
        else inner_fwd_big_kernel<float, 16, false><<<grid, nt, smem, st>>>(view_of(left), view_of(right), (float*)out, g, dchp, xtiles, vec8, none);
        return finish_launch("rsm_inner_fwd(8xDT)");
      }
    }
  }
  return RSM_DISPATCH_IO(in_dtype, out_dtype, Tin, Tout, [&]() -> int {
    return launch_fwd<Tin, Tout, LAYOUT_NDHW>(left, right, out, N, g, st, "rsm_inner_fwd");
  });
}

static int inner_bwd(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright, int64_t N, int64_t C,
                     int64_t H, int64_t W, int64_t D, int reduce, int in_dtype, int out_dtype, int device, void* stream,
                     unsigned long long* prof) {
  CorrGeom g;
  if (int rc = make_geom(N, C, H, W, D, 1, reduce == RSM_REDUCE_MEAN, false, g)) return rc;
  if (N * C * H * W == 0) return RSM_OK;
  if (!left.data || !right.data || (D > 0 && !gout)) return RSM_ERR_NULL_POINTER;
  if (!gleft && !gright) return RSM_OK;
  RSM_COMMON_CHECKS(in_dtype)
  if (!valid_dtype(out_dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  {   // 16-bit tensors, D <= 64: band matrices + tcgen05 (rsm_corr_bwd_tc.cu)
    const int rc = launch_inner_bwd_tc(gout, left, right, gleft, gright, N, C, H, W, D, reduce == RSM_REDUCE_MEAN, in_dtype,
                                       out_dtype, st, prof);
    if (rc != RSM_ERR_UNSUPPORTED_CONFIG) return rc;
  }
  return RSM_DISPATCH_IO(in_dtype, out_dtype, Tin, Tout, [&]() -> int {
    return launch_bwd<Tin, Tout, LAYOUT_NDHW>(gout, left, right, gleft, gright, N, g, st, "rsm_inner_bwd");
  });
}

extern "C" int rsm_inner_bwd(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                             int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int reduce,
                             int in_dtype, int out_dtype, int device, void* stream) {
  return inner_bwd(gout, left, right, gleft, gright, N, C, H, W, D, reduce, in_dtype, out_dtype, device, stream, nullptr);
}

// diagnostic twin: `prof` = 16 zero-initialised uint64 on the device; the tcgen05 adjoint kernel adds clock64 cycles
// per warp role (other paths leave it untouched)
extern "C" int rsm_inner_bwd_profile(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                                     int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int reduce,
                                     int in_dtype, int out_dtype, int device, void* stream, uint64_t* prof) {
  if (!prof) return RSM_ERR_NULL_POINTER;
  return inner_bwd(gout, left, right, gleft, gright, N, C, H, W, D, reduce, in_dtype, out_dtype, device, stream,
                   reinterpret_cast<unsigned long long*>(prof));
}

extern "C" int rsm_groupwise_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                                 int64_t W, int64_t D, int64_t G, int in_dtype, int out_dtype, int device,
                                 void* stream) {
  CorrGeom g;
  if (int rc = make_geom(N, C, H, W, D, G, 1, false, g)) return rc;
  if (N * G * H * W * D == 0) return RSM_OK;
  if (!out || (C > 0 && (!left.data || !right.data))) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(in_dtype)
  if (!valid_dtype(out_dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (!aligned_to(out, dtype_size(out_dtype))) return RSM_ERR_MISALIGNED;
  {   // groups of <= 32 channels, rows TMA can address: the persistent TMA-fed kernel (rsm_groupwise_tma.cu)
    const int rc = launch_groupwise_tma(left, right, out, N, C, H, W, D, G, in_dtype, out_dtype, st);
    if (rc != RSM_ERR_UNSUPPORTED_CONFIG) return rc;
  }
  return RSM_DISPATCH_IO(in_dtype, out_dtype, Tin, Tout, [&]() -> int {
    return launch_fwd<Tin, Tout, LAYOUT_NGHWD>(left, right, out, N, g, st, "rsm_groupwise_fwd");
  });
}

extern "C" int rsm_groupwise_bwd(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                                 int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G,
                                 int in_dtype, int out_dtype, int device, void* stream) {
  CorrGeom g;
  if (int rc = make_geom(N, C, H, W, D, G, 1, false, g)) return rc;
  if (N * C * H * W == 0) return RSM_OK;
  if (!left.data || !right.data || (D > 0 && !gout)) return RSM_ERR_NULL_POINTER;
  if (!gleft && !gright) return RSM_OK;
  RSM_COMMON_CHECKS(in_dtype)
  if (!valid_dtype(out_dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  return RSM_DISPATCH_IO(in_dtype, out_dtype, Tin, Tout, [&]() -> int {
    return launch_bwd<Tin, Tout, LAYOUT_NGHWD>(gout, left, right, gleft, gright, N, g, st, "rsm_groupwise_bwd");
  });
}

static int inner_regress_fwd(rsm_feat left, rsm_feat right, int64_t N, int64_t C, int64_t H, int64_t W, int64_t D,
                             int reduce, int in_dtype, rsm_regress_out out, int device, void* stream,
                             unsigned long long* prof) {
  CorrGeom g;
  if (D <= 0) return RSM_ERR_INVALID_SHAPE;
  if (out.soft && out.expect && out.soft != (void*)out.expect) return RSM_ERR_UNSUPPORTED_CONFIG;   // both are the fp32 plane here
  if (!out.soft) out.soft = out.expect;
  if (int rc = make_geom(N, C, H, W, D, 1, reduce == RSM_REDUCE_MEAN, true, g)) return rc;
  if (N * H * W == 0) return RSM_OK;
  if (C > 0 && (!left.data || !right.data)) return RSM_ERR_NULL_POINTER;
  if (NTX * g.ntd > 1024) return RSM_ERR_UNSUPPORTED_CONFIG;  // D <= 512
  RSM_COMMON_CHECKS(in_dtype)
  {   // 16-bit features (C % 16 == 0, C <= 128, D <= 384, TMA-addressable views): tcgen05 accumulators reduced in TMEM
    const int rc = launch_inner_regress_rows(left, right, N, C, H, W, D, reduce == RSM_REDUCE_MEAN, in_dtype, out, st, prof);
    if (rc != RSM_ERR_UNSUPPORTED_CONFIG) return rc;
  }
  // fp32 features, all disparities in one 64-wide chunk, rows on 16-byte boundaries: the 8xDT-tile kernel with
  // the regression scan as its epilogue
  {
    auto v4 = [&](const rsm_feat& f) {
      return f.stride_w == 1 && f.stride_n % 4 == 0 && f.stride_c % 4 == 0 && f.stride_h % 4 == 0 && aligned_to(f.data, 16);
    };
    if (in_dtype == RSM_F32 && C >= 16 && D >= 16 && D <= 16 * BG_MAXNTD && W % 4 == 0 && v4(left) && v4(right)) {
      static const int kDT[5] = {8, 12, 8, 12, 16}, kNTD[5] = {2, 2, 4, 4, 4};   // chunks 16, 24, 32, 48, 64
      int pick = 4;
      for (int k = 0; k < 5; ++k)
        if (kDT[k] * kNTD[k] >= D) { pick = k; break; }
      const int dt = kDT[pick], dchp = dt * kNTD[pick];
      const int xtiles = (int)ceil_div(W, BG_TX);
      const int64_t bxb = N * H * xtiles;
      if (grid_ok(bxb)) {
        const size_t smemb = (size_t)BG_CK * (2 * BG_TX + dchp) * sizeof(float);   // >= BG_TX * (dchp + 1) floats
        const RegressOutPtrs ro{(float*)out.soft, out.argmin, out.argmax, out.lse};
        const unsigned nt = BG_NTX * kNTD[pick];
        if (dt == 8) inner_fwd_big_kernel<float, 8, true><<<(unsigned)bxb, nt, smemb, st>>>(view_of(left), view_of(right), nullptr, g, dchp, xtiles, 0, ro);
        else if (dt == 12) inner_fwd_big_kernel<float, 12, true><<<(unsigned)bxb, nt, smemb, st>>>(view_of(left), view_of(right), nullptr, g, dchp, xtiles, 0, ro);
        else inner_fwd_big_kernel<float, 16, true><<<(unsigned)bxb, nt, smemb, st>>>(view_of(left), view_of(right), nullptr, g, dchp, xtiles, 0, ro);
        return finish_launch("rsm_inner_regress_fwd(8xDT)");
      }
    }
  }
  const int64_t bx = N * g.H * g.xtiles;
  if (!grid_ok(bx)) return RSM_ERR_INVALID_SHAPE;
  const size_t smem = (size_t)(CKMAX * TX + CKMAX * (TX + g.dchp) + TX * (g.dchp + 1)) * sizeof(float);
  return RSM_DISPATCH_DTYPE(in_dtype, Tin, [&]() -> int {
    auto k = (NTX * g.ntd <= 512) ? inner_regress_fwd_kernel<Tin, 512> : inner_regress_fwd_kernel<Tin, 1024>;
    if (smem > 48 * 1024) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return finish_launch("rsm_inner_regress_fwd(attr)");
    }
    k<<<(unsigned)bx, NTX * g.ntd, smem, st>>>(view_of(left), view_of(right), (float*)out.soft, out.argmin,
                                               out.argmax, out.lse, g);
    return finish_launch("rsm_inner_regress_fwd");
  });
}

extern "C" int rsm_inner_regress_fwd(rsm_feat left, rsm_feat right, int64_t N, int64_t C, int64_t H,
                                     int64_t W, int64_t D, int reduce, int in_dtype, rsm_regress_out out,
                                     int device, void* stream) {
  return inner_regress_fwd(left, right, N, C, H, W, D, reduce, in_dtype, out, device, stream, nullptr);
}

// diagnostic twin: `prof` = 16 zero-initialised uint64 on the device; the row-streaming tcgen05 kernel adds clock64
// cycles per warp role (other paths leave it untouched)
extern "C" int rsm_inner_regress_fwd_profile(rsm_feat left, rsm_feat right, int64_t N, int64_t C, int64_t H,
                                             int64_t W, int64_t D, int reduce, int in_dtype, rsm_regress_out out,
                                             int device, void* stream, uint64_t* prof) {
  if (!prof) return RSM_ERR_NULL_POINTER;
  return inner_regress_fwd(left, right, N, C, H, W, D, reduce, in_dtype, out, device, stream,
                           reinterpret_cast<unsigned long long*>(prof));
}
