// Inner-product / correlation volume on the 5th-gen tensor cores (tcgen05) for 16-bit features.
//
// Per epipolar row the correlation is the band  0 <= x - x' < D  of the W x W product
// P[x, x'] = sum_c L[c, x] * R[c, x']  (the reference's own einsum hint, cost_volume/inner_product.py:33-34).
// A CTA owns TM = 128 left pixels x0.. of one (n, y) and a chunk of DCH <= 128 disparities dc0..:
//     D_tmem[r, j] = sum_c L[c, x0 + r] * R[c, xr0 + j],   xr0 = x0 - dc0 - DCH,  j in [0, 128 + DCH)
// is ONE tcgen05.mma per 16 channels (M = 128, N = 128 + DCH, K = 16, both operands MN-major in shared
// memory, fp32 accumulators in TMEM, issued by one thread).  The wanted value for disparity
// dc0 + dl of pixel x0 + r sits at column j = r + DCH - dl: a diagonal band.  Each epilogue warp
// (TMEM lanes 32w..32w+31) pulls the DCH + 32 columns that cover its lanes with tcgen05.ld, parks
// them in a padded shared-memory row per lane, and reads them back skewed so that for every
// disparity the 32 lanes store 32 consecutive x of the (N,D,H,W) volume.
//
// Operand staging is done with ordinary vector loads (the features may be strided views and the
// right window needs zero fill on both sides); the canonical no-swizzle MN-major core-matrix
// layout is written directly:  addr(x, c) = ((c/8) * (T/8) + x/8) * 128 + (c%8) * 16 + (x%8) * 2.
#include <stdlib.h>

#include "rsm_common.cuh"

namespace rsm {

constexpr int TC_TM = 128;   // UMMA M: left pixels per CTA
constexpr int TC_KC = 64;    // channels per shared-memory stage

struct TcGeom {
  int C, H, W, D;
  int dch;      // disparities per CTA chunk (multiple of 16, <= 128)
  int ncol;     // UMMA N = TC_TM + dch
  int pitch;    // floats per lane row of the skew buffer
  int xtiles;   // ceil(W / TC_TM)
  int mean, pow2;
  int fmt;      // 0 = fp16, 1 = bf16 (UMMA a/b format)
  int tmem_cols;
  int smem_main;  // bytes of max(operand buffers, skew buffer)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// wait for completion of the given phase of an mbarrier; bounded so a protocol bug cannot hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t phase) {
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(phase)
        : "memory");
    if (ok) return true;
  }
  return false;
}

// ---- stage nch channels of one operand: xs = first x of the tile, nxg = x-groups of 8.
// thread -> (channel inside its K-group: 8 lanes write 128 contiguous bytes, x-group lane); all
// K-groups of one x-group are loaded before any is stored (up to 8 independent 16-byte loads in flight)
template <typename Tin>
__device__ __forceinline__ uint4 load_chunk_slow(const Tin* __restrict__ src, int x, int W, int64_t sw) {
  union { uint4 u; Tin e[8]; } tmp;
  tmp.u = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (x + i >= 0 && x + i < W) tmp.e[i] = __ldg(src + (int64_t)(x + i) * sw);
  return tmp.u;
}

template <typename Tin>
__device__ __forceinline__ void stage_operand(const FeatView& F, int64_t n, int y, int c0, int nch, int xs, int nxg, int W,
                                              unsigned char* dst, bool fast) {
  const int cl = threadIdx.x & 7, xl = threadIdx.x >> 3, nxl = blockDim.x >> 3;
  const int ncg = nch >> 3;
  const Tin* __restrict__ base =
      reinterpret_cast<const Tin*>(F.data) + n * F.sn + (int64_t)y * F.sh + (int64_t)(c0 + cl) * F.sc;
  for (int xg = xl; xg < nxg; xg += nxl) {
    const int x = xs + 8 * xg;
    const bool inside = fast && x >= 0 && x + 8 <= W;
    const bool empty = x + 8 <= 0 || x >= W;
    for (int cg0 = 0; cg0 < ncg; cg0 += 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (cg0 + u < ncg && !empty) {
          const Tin* src = base + (int64_t)(8 * (cg0 + u)) * F.sc;
          v[u] = inside ? __ldg(reinterpret_cast<const uint4*>(src + x)) : load_chunk_slow<Tin>(src, x, W, F.sw);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (cg0 + u < ncg)
          *reinterpret_cast<uint4*>(dst + ((size_t)((cg0 + u) * nxg + xg) * 8 + cl) * 16) = v[u];
    }
  }
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(128)
inner_tc_fwd_kernel(FeatView L, FeatView R, Tout* __restrict__ out, TcGeom g, int fast) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                                   // TC_KC * TC_TM * 2 bytes
  unsigned char* sB = sA + TC_KC * TC_TM * 2;                      // TC_KC * ncol * 2 bytes
  float* skew = reinterpret_cast<float*>(smem_raw);                // 128 * pitch floats, ALIASES the operand
                                                                   // buffers (free once the last MMA committed)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + g.smem_main);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % g.xtiles); bid /= g.xtiles;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * TC_TM;
  const int dc0 = blockIdx.y * g.dch;
  const int xr0 = x0 - dc0 - g.dch;

  // ---- one-time setup: TMEM allocation (warp 0), mbarrier (thread 0)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor: D = f32, A/B = fmt, both MN-major, N = ncol, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)g.fmt << 7) | ((uint32_t)g.fmt << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(g.ncol >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
  const uint32_t sbo = 128;                              // next 8-pixel group along M / N
  const uint32_t lboA = (TC_TM / 8) * 128;               // next 8-channel group along K
  const uint32_t lboB = (uint32_t)(g.ncol / 8) * 128;

  uint32_t phase = 0;
  bool ok = true;
  for (int c0 = 0; c0 < g.C; c0 += TC_KC) {
    const int nch = min(TC_KC, g.C - c0);
    stage_operand<Tin>(L, n, y, c0, nch, x0, TC_TM / 8, g.W, sA, fast);
    stage_operand<Tin>(R, n, y, c0, nch, xr0, g.ncol / 8, g.W, sB, fast);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < nch / 16; ++ks) {
        // descriptor fields (cute make_umma_desc<Major::MN>, SWIZZLE_NONE): SBO = stride between 8-element
        // groups along M/N, LBO = stride between 8-row groups along K
        // (verified on B200 against the oracle; the swapped assignment produces garbage)
        const uint64_t adesc = umma_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo);
        const uint64_t bdesc = umma_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo);
        umma_f16(tmem_base, adesc, bdesc, idesc, (c0 > 0 || ks > 0) ? 1u : 0u);
      }
      // completion of all MMAs issued so far -> mbarrier (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                   : "memory");
    }
    ok = mbar_wait(smem_u32(mbar), phase) && ok;   // operands consumed: smem may be restaged, TMEM is current
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: lane r = 32*warp + lane owns pixel x0 + r; columns [32*warp, 32*warp + dch + 32)
  float* row = skew + (size_t)(32 * warp + lane) * g.pitch;
  const int ncw = g.dch + 32;
  for (int cb = 0; cb < ncw; cb += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(32 * warp + cb), r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *reinterpret_cast<uint4*>(row + cb + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
  }
  __syncwarp();
  const int x = x0 + 32 * warp + lane;
  const float inv = 1.f / (float)g.C, cnt = (float)g.C;
  if (x < g.W) {
    const int dmax = min(g.dch, g.D - dc0);
    Tout* __restrict__ o = out + (((int64_t)n * g.D + dc0) * g.H + y) * g.W + x;
    const int64_t dstride = (int64_t)g.H * g.W;
    for (int dl = 0; dl < dmax; ++dl) {
      float v = ok ? row[lane + g.dch - dl] : __int_as_float(0x7fc00000);
      if (g.mean) v = g.pow2 ? v * inv : v / cnt;
      if (x < dc0 + dl) v = 0.f;                       // the reference leaves zeros where x < d
      o[dl * dstride] = from_f<Tout>(v);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when the tensor-core path does not apply (caller falls back to SIMT)
template <typename Tin, typename Tout>
static int launch_inner_tc_typed(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H,
                                 int64_t W, int64_t D, int mean, int fmt, cudaStream_t st) {
  TcGeom g;
  g.C = (int)C; g.H = (int)H; g.W = (int)W; g.D = (int)D;
  const int d16 = (int)((D + 15) / 16 * 16);
  g.dch = d16 < 128 ? d16 : 128;
  g.ncol = TC_TM + g.dch;
  int p = g.dch + 32;                    // pitch: >= dch + 32, multiple of 4 with an odd quotient (conflict-free
  p = (p + 3) / 4 * 4;                   // 128-bit row writes and conflict-free skewed 32-bit reads)
  if ((p / 4) % 2 == 0) p += 4;
  g.pitch = p;
  g.xtiles = (int)ceil_div(W, TC_TM);
  g.mean = mean;
  g.pow2 = (C & (C - 1)) == 0;
  g.fmt = fmt;
  g.tmem_cols = g.ncol <= 128 ? 128 : 256;
  const int64_t bx = N * H * g.xtiles, by = ceil_div(D, g.dch);
  if (bx <= 0 || bx > 2147483647LL || by > 65535) return RSM_ERR_INVALID_SHAPE;
  const size_t ops = (size_t)TC_KC * TC_TM * 2 + (size_t)TC_KC * g.ncol * 2, skw = (size_t)TC_TM * g.pitch * 4;
  g.smem_main = (int)(((ops > skw ? ops : skw) + 15) / 16 * 16);
  const size_t smem = (size_t)g.smem_main + 16;
  auto k = inner_tc_fwd_kernel<Tin, Tout>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return finish_launch("rsm_inner_fwd(tc attr)");
  auto vec_ok = [&](const rsm_feat& f) {
    return f.stride_w == 1 && f.stride_n % 8 == 0 && f.stride_c % 8 == 0 && f.stride_h % 8 == 0 && aligned_to(f.data, 16);
  };
  const int fast = vec_ok(left) && vec_ok(right);   // 16-byte chunks start at multiples of 8 elements
  k<<<dim3((unsigned)bx, (unsigned)by), 128, smem, st>>>(view_of(left), view_of(right), (Tout*)out, g, fast);
  return finish_launch("rsm_inner_fwd(tcgen05)");
}

int launch_inner_tc(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                    int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st) {
  if (in_dtype == RSM_F32 || C % 16 != 0 || C <= 0 || D <= 0) return RSM_ERR_UNSUPPORTED_CONFIG;
  if (in_dtype == RSM_F16) {
    if (out_dtype == RSM_F32) return launch_inner_tc_typed<__half, float>(left, right, out, N, C, H, W, D, mean, 0, st);
    return launch_inner_tc_typed<__half, __half>(left, right, out, N, C, H, W, D, mean, 0, st);
  }
  if (out_dtype == RSM_F32) return launch_inner_tc_typed<__nv_bfloat16, float>(left, right, out, N, C, H, W, D, mean, 1, st);
  return launch_inner_tc_typed<__nv_bfloat16, __nv_bfloat16>(left, right, out, N, C, H, W, D, mean, 1, st);
}

}  // namespace rsm
