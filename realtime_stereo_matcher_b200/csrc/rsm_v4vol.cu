// MobileStereoNetV4's per-disparity learned cost volume (model/mobile_stereo_net_v4.py:443-458) as ONE fused op in
// eval mode: for every disparity i the reference interweaves featL[..., i:] with featR[..., :-i], runs three strided
// Conv3d (+BatchNorm3d+ReLU) that collapse the 64-deep interleaved axis (:317-333), a 1x1 conv (+BN+ReLU, :335) and
// writes volume[:, 0, i, :, i:].  48 Python iterations, 72 % of the v4 forward.
//
// Depth stride = kernel depth in every Conv3d, so each layer is a 3x3 2-D convolution over independent depth blocks
// with shared weights; BatchNorm (eval) folds into weights + bias.  Three kernels:
//
//  K1 v4_premap_kernel   Layer 1 is linear in its input and its input is (left, right shifted by d) interleaved, so
//                        conv1(d)[x] = PL[x] + PR[x - d] with PL / PR the left-only / right-only halves of the 3x3x8
//                        kernel applied ONCE per pair (not per disparity): 48x fewer layer-1 flops.  The reference
//                        convolves the CROPPED tensor (x >= d), i.e. zero padding at x = d-1 and at the right end of
//                        the cropped right image; two edge maps (EL: the dx=-1 column of PL, ER: the dx=+1 column of
//                        PR) are subtracted at x == d and x == W-1 to reproduce that exactly.  Maps are stored as
//                        16-bit (chunk of 8 channels innermost): [b][16 chunks][H][W][8].
//  K2 v4_conv_kernel<32,GEN>  layer 2 (16ch x 4 depths x 3x3 -> 32, K = 576) as an implicit GEMM on tcgen05: a
//                        persistent CTA walks a strip (pair b, disparity d, depth block j, 128-pixel column) top to
//                        bottom; producer warps build each input row ONCE in shared memory (relu(PL + PR - edges),
//                        x < d zeroed, 1-pixel halo) in the K-major no-swizzle core-matrix layout
//                        [8-channel chunk][pixel][8] (16 B per pixel per chunk), so the three dx taps are the same row
//                        at a 16-byte offset and the three dy taps are three slots of a 6-row ring: every input value
//                        is staged once and read by nine taps.  One thread issues 36 tcgen05.mma (M=128 pixels, N=32,
//                        K=16) per output row into one of four TMEM accumulators; epilogue warps add the folded bias,
//                        ReLU, re-zero x < d and store 16-bit activations [d][b][8 chunks][H][W][8].
//  K3 v4_conv_kernel<16,TMA>  layer 3 (32ch x 2 depths x 3x3 -> 16, K = 576) the same way with the rows arriving by
//                        TMA (one 5-D box of 130 pixels x 8 chunks per row, out-of-image pixels zero-filled by the
//                        TMA unit), fused with the 1x1 conv + BN + ReLU and the x < d mask; writes volume[b, d, y, x].
//
// Operands are fp16 (fp32 / fp16 features; 11-bit significand = TF32's, which is what cuDNN uses for fp32 convs by
// default; values saturate at +-65504) or bf16 (bf16 features); accumulation is fp32.
#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int V4_TM = 128;                 // pixels per tile = UMMA M
constexpr int V4_PX = V4_TM + 2;           // staged pixels per row: 1-pixel halo on both sides
constexpr int V4_KCH = 8;                  // 8-channel chunks per staged row (64 input channels per tap)
constexpr int V4_ROW_BYTES = V4_KCH * V4_PX * 16;   // 16640
constexpr int V4_NROWS = 6;                // ring of staged input rows
constexpr int V4_NACC = 4;                 // TMEM accumulator buffers
constexpr int V4_C = 32;                   // feature channels per image (2C = 64 interleaved = 8 depth blocks x 8)
constexpr int V4_C1 = 16, V4_C2 = 32, V4_C3 = 16;   // Conv3d output channels

// ------------------------------------------------------------------------------------------------ conversions
template <typename T16> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  uint32_t r;   // saturating: +-inf would poison the softmax downstream, the reference's fp16 convs clamp nothing but
                // also never see values this large after BatchNorm
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <typename T16> __device__ __forceinline__ float2 unpack2h(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2h<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2h<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T16> __device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = unpack2h<T16>(v.x), b = unpack2h<T16>(v.y), c = unpack2h<T16>(v.z), d = unpack2h<T16>(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T16> __device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2<T16>(f[0], f[1]), pack2<T16>(f[2], f[3]), pack2<T16>(f[4], f[5]), pack2<T16>(f[6], f[7]));
}

// relu(a + b) on two packed 16-bit values, clamped to the largest finite value (the sum of two saturated maps may
// overflow fp16).  A correctly rounded 16-bit add of two 16-bit values equals their exact sum rounded once, i.e. the
// same result as adding in fp32 and rounding -- at a quarter of the instructions.
template <typename T16> __device__ __forceinline__ uint32_t add_relu2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t add_relu2<__half>(uint32_t a, uint32_t b) {
  const __half2 z = __float2half2_rn(0.f), m = __float2half2_rn(65504.f);
  __half2 s = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  s = __hmin2(__hmax2(s, z), m);
  return *reinterpret_cast<const uint32_t*>(&s);
}
template <> __device__ __forceinline__ uint32_t add_relu2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  const __nv_bfloat162 z = __float2bfloat162_rn(0.f), m = __float2bfloat162_rn(3.3895314e38f);
  __nv_bfloat162 s = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  s = __hmin2(__hmax2(s, z), m);
  return *reinterpret_cast<const uint32_t*>(&s);
}

// ================================================================================================ K1: layer-1 maps
// grid (ceil(W/32), H, 2B: left / right maps), 256 threads: warp = depth block k (interleaved channels 8k..8k+7 = left 4k..4k+3 at even
// depths, right 4k..4k+3 at odd depths), lane = pixel.  w1 (16, 8, 3, 3) fp32 with BatchNorm folded, t1 (16) the
// folded bias (goes into PL).  Map channel = k*16 + o, stored [b][chunk = 2k + o/8][y][x][o%8].
template <typename Tin, typename T16>
__global__ void __launch_bounds__(256)
v4_premap_kernel(FeatView L, FeatView R, const float* __restrict__ w1, const float* __restrict__ t1,
                 uint4* __restrict__ PL, uint4* __restrict__ PR, uint4* __restrict__ EL, uint4* __restrict__ ER, int H, int W) {
  __shared__ float sw[4 * 9 * 16];     // this side's depths: [cl][tap][o]
  __shared__ float st[16];
  const int side = blockIdx.z & 1, b = blockIdx.z >> 1;     // 0: left maps (PL, EL), 1: right maps (PR, ER)
  for (int i = threadIdx.x; i < 4 * 9 * 16; i += 256) {
    const int o = i & 15, tap = (i >> 4) % 9, cl = i / 144;
    sw[i] = w1[(o * 8 + 2 * cl + side) * 9 + tap];
  }
  if (threadIdx.x < 16) st[threadIdx.x] = side == 0 ? t1[threadIdx.x] : 0.f;   // the folded bias goes into PL
  __syncthreads();
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y;
  if (x >= W) return;
  float acc[16], edg[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) { acc[o] = st[o]; edg[o] = 0.f; }
  const FeatView& F = side == 0 ? L : R;
  const Tin* __restrict__ pf = reinterpret_cast<const Tin*>(F.data) + (int64_t)b * F.sn + (int64_t)(4 * k) * F.sc;
  const int edge_dx = side == 0 ? 0 : 2;                    // EL: the column that reads x - 1; ER: the one that reads x + 1
#pragma unroll
  for (int cl = 0; cl < 4; ++cl) {
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        const float v = (xx >= 0 && xx < W) ? to_f(__ldg(pf + (int64_t)cl * F.sc + (int64_t)yy * F.sh + (int64_t)xx * F.sw)) : 0.f;
        const float4* wp = reinterpret_cast<const float4*>(sw + (cl * 9 + dy * 3 + dx) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = wp[q];
          const float wa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[4 * q + i] = fmaf(wa[i], v, acc[4 * q + i]);
            if (dx == edge_dx) edg[4 * q + i] = fmaf(wa[i], v, edg[4 * q + i]);
          }
        }
      }
    }
  }
  uint4* __restrict__ P = side == 0 ? PL : PR;
  uint4* __restrict__ E = side == 0 ? EL : ER;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t o = (((int64_t)b * 16 + 2 * k + half) * H + y) * W + x;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = acc[8 * half + i];
    P[o] = pack8<T16>(v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = edg[8 * half + i];
    E[o] = pack8<T16>(v);
  }
}

// ======================================================================================= K2 / K3: implicit GEMM
struct V4Geom {
  int B, H, W, D;
  int xtiles;            // ceil(W / 128)
  int nj;                // depth blocks per (b, d): 2 for layer 2, 1 for layer 3
  int strips;            // B * D * nj * xtiles
  int fmt;               // 0 = fp16, 1 = bf16
};

struct V4Strip { int b, d, j, x0; };
__device__ __forceinline__ V4Strip v4_strip(int s, const V4Geom& g) {
  V4Strip r;
  const int xt = s % g.xtiles; s /= g.xtiles;
  r.j = s % g.nj; s /= g.nj;
  r.d = s % g.D; r.b = s / g.D;
  r.x0 = xt * V4_TM;
  return r;
}

__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// NOUT: output channels (UMMA N).  GEN: rows generated from the layer-1 maps (layer 2) / rows by TMA (layer 3).
// Warps: 0-3 epilogue (TMEM lane quadrant = warp), 4 UMMA issuer, 5.. producers (GEN: 8 warps in two groups that
// alternate rows, so two rows of L2 loads are in flight; TMA: one lane of warp 5).
template <int NOUT, bool GEN, typename T16, typename Tout>
__global__ void __launch_bounds__(GEN ? 13 * 32 : 6 * 32, 1)
v4_conv_kernel(const uint4* __restrict__ PL, const uint4* __restrict__ PR, const uint4* __restrict__ EL,
               const uint4* __restrict__ ER, const __grid_constant__ CUtensorMap tmIn, const uint4* __restrict__ wpacked,
               const float* __restrict__ bias, const float* __restrict__ w11, const float* __restrict__ t11,
               uint4* __restrict__ act_out, Tout* __restrict__ vol, V4Geom g, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int WBYTES = 9 * V4_KCH * NOUT * 16;
  unsigned char* wsm = ring + V4_NROWS * V4_ROW_BYTES;
  float* fsm = reinterpret_cast<float*>(wsm + WBYTES);          // bias[NOUT], w11[16], t11
  uint64_t* bars = reinterpret_cast<uint64_t*>(fsm + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * V4_NROWS + 2 * V4_NACC);
  const uint32_t full = smem_u32(bars), empty = smem_u32(bars + V4_NROWS), accfull = smem_u32(bars + 2 * V4_NROWS),
                 accempty = smem_u32(bars + 2 * V4_NROWS + V4_NACC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = V4_NACC * NOUT < 32 ? 32 : V4_NACC * NOUT;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < V4_NROWS; ++i) {
      mbar_init(full + 8 * i, GEN ? 128 : 1);     // a producer group's 128 threads / the TMA lane's expect_tx
      mbar_init(empty + 8 * i, 1);                // one UMMA commit
    }
    for (int i = 0; i < V4_NACC; ++i) {
      mbar_init(accfull + 8 * i, 1);
      mbar_init(accempty + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights (already in the [tap][chunk][co][8] core-matrix layout) and epilogue constants -> shared memory
  for (int i = threadIdx.x; i < WBYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(wsm)[i] = wpacked[i];
  if (threadIdx.x < NOUT) fsm[threadIdx.x] = bias[threadIdx.x];
  if (!GEN && threadIdx.x < 16) fsm[32 + threadIdx.x] = w11[threadIdx.x];
  if (!GEN && threadIdx.x == 0) fsm[48] = t11[0];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int H = g.H, W = g.W;

  if (warp == 4) {
    // ================================================================================ UMMA issuer
    // The whole warp runs the loop converged (every value below is warp-uniform, so the compiler keeps it in uniform
    // registers) and ONE elected lane issues.  The first version branched on lane == 0 and rebuilt both 64-bit
    // descriptors per MMA: 13 dependent instructions, ~68 cycles per tcgen05.mma measured with clock64 -- the issue
    // stream, not the tensor pipe, set the pace (tensor pipe 10-15 % busy).  Now: descriptor high words are constants,
    // low words are a base plus an immediate.
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | ((uint32_t)g.fmt << 7) | ((uint32_t)g.fmt << 10) | ((uint32_t)(NOUT >> 3) << 17) |
                           ((uint32_t)(V4_TM >> 4) << 24);        // K-major A and B, fp32 accumulate
    // K-major, SWIZZLE_NONE: LBO = stride between the two 8-channel core matrices of a K = 16 step (the next chunk),
    // SBO = stride between 8-row groups along M / N (8 pixels or 8 output channels x 16 B) -- verified on B200
    // against the reference loop (the swapped assignment produces garbage)
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
    const uint32_t alo0 = ((smem_u32(ring) >> 4) & 0x3FFF) | ((uint32_t)(V4_PX * 16 >> 4) << 16);
    const uint32_t blo0 = ((smem_u32(wsm) >> 4) & 0x3FFF) | ((uint32_t)(NOUT * 16 >> 4) << 16);
    uint32_t rid0 = 0;          // row id of this strip's row 0 (ids run on across the CTA's strips)
    uint32_t waited = 0;        // rows [0, waited) of the id sequence have been waited for
    uint32_t it = 0;            // output rows issued so far
    long long c_full = 0, c_acc = 0;
    const long long c_beg = prof ? clock64() : 0;
    for (int s = blockIdx.x; s < g.strips; s += gridDim.x, rid0 += (uint32_t)H) {
      for (int y = 0; y < H; ++y, ++it) {
        const uint32_t need = rid0 + (uint32_t)min(y + 1, H - 1) + 1;       // rows up to y+1 must have landed
        const long long c0 = prof ? clock64() : 0;
        for (; waited < need; ++waited) mbar_wait(full + 8 * (waited % V4_NROWS), (waited / V4_NROWS) & 1);
        const long long c1 = prof ? clock64() : 0;
        const uint32_t ab = it % V4_NACC;
        mbar_wait(accempty + 8 * ab, ((it / V4_NACC) & 1) ^ 1);             // epilogue drained this accumulator
        if (prof) { c_full += c1 - c0; c_acc += clock64() - c1; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t td = tmem_base + ab * NOUT;
        if (leader) {
          uint32_t acc = 0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= H) continue;                                  // zero padding rows: no MMAs at all
            const uint32_t alo = alo0 + ((rid0 + (uint32_t)yy) % V4_NROWS) * (V4_ROW_BYTES >> 4);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int ks = 0; ks < V4_KCH / 2; ++ks) {                        // K = 16 per UMMA: two 8-channel chunks
                const uint32_t a = alo + (uint32_t)(dx + ks * 2 * V4_PX);                               // 16-byte units
                const uint32_t b = blo0 + (uint32_t)(((dy * 3 + dx) * V4_KCH + 2 * ks) * NOUT);
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                    "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                    ::"r"(td), "r"(a), "r"(b), "r"(DESC_HI), "r"(idesc), "r"(acc)
                    : "memory");
                acc = 1;
              }
            }
          }
          umma_commit(accfull + 8 * ab);                                       // accumulator ready for the epilogue
          if (y >= 1) umma_commit(empty + 8 * ((rid0 + (uint32_t)y - 1) % V4_NROWS));   // row y-1 is no longer needed
          if (y == H - 1) umma_commit(empty + 8 * ((rid0 + (uint32_t)y) % V4_NROWS));
        }
        __syncwarp();
      }
    }
    if (prof && leader) {
      atomicAdd(prof + 0, (unsigned long long)c_full);
      atomicAdd(prof + 1, (unsigned long long)c_acc);
      atomicAdd(prof + 2, (unsigned long long)(clock64() - c_beg));
      atomicAdd(prof + 11, (unsigned long long)it);
    }
  } else if (warp >= 5) {
    // ================================================================================== producers
    if constexpr (GEN) {
      const int grp = (warp - 5) >> 2, t = (threadIdx.x - 5 * 32) & 127;
      const int plane = H * W;                                                // (B * 16 * H * W < 2^31: checked on the host)
      constexpr int NIT = (V4_KCH * V4_PX + 127) / 128;                       // 9 items of 16 B per thread and row
      const bool rec = prof && t == 0;
      long long c_wait = 0, c_load = 0;
      const long long c_beg = rec ? clock64() : 0;
      uint32_t rid0 = 0;
      for (int s = blockIdx.x; s < g.strips; s += gridDim.x, rid0 += (uint32_t)H) {
        const V4Strip st = v4_strip(s, g);
        // per strip: where each of this thread's items lives in the maps (row 0), whether it is inside the cropped
        // image (x >= d, x < W: else the staged value is zero) and whether it is one of the two edge columns
        int off[NIT];
        uint32_t live = 0, edge = 0;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int item = t + 128 * i, c = item / V4_PX, p = item - c * V4_PX, x = st.x0 - 1 + p;
          const bool ok = item < V4_KCH * V4_PX && x >= st.d && x < W;
          off[i] = (st.b * 16 + 8 * st.j + c) * plane + x;
          live |= (uint32_t)ok << i;
          edge |= (uint32_t)(ok && (x == st.d || x == W - 1)) << i;
        }
        for (int y = 0; y < H; ++y) {
          const uint32_t rid = rid0 + (uint32_t)y;
          if ((int)(rid & 1) != grp) continue;
          const uint32_t slot = rid % V4_NROWS;
          const long long c0 = rec ? clock64() : 0;
          mbar_wait(empty + 8 * slot, ((rid / V4_NROWS) & 1) ^ 1);           // UMMAs that read this slot have completed
          const long long c1 = rec ? clock64() : 0;
          uint4* dst = reinterpret_cast<uint4*>(ring + slot * V4_ROW_BYTES) + t;
          const int ro = y * W;
          uint4 vl[NIT], vr[NIT];
#pragma unroll
          for (int i = 0; i < NIT; ++i) {                                     // all of the row's loads in flight first
            const bool ok = (live >> i) & 1;
            vl[i] = ok ? __ldg(PL + off[i] + ro) : make_uint4(0u, 0u, 0u, 0u);
            vr[i] = ok ? __ldg(PR + off[i] + ro - st.d) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            if (i == NIT - 1 && t + 128 * i >= V4_KCH * V4_PX) break;
            uint4 v;
            if ((edge >> i) & 1) {                                            // the cropped tensor's zero padding (rare)
              float a[8], r[8];
              unpack8<T16>(vl[i], a);
              unpack8<T16>(vr[i], r);
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] += r[e];
              const int x = st.x0 - 1 + (t + 128 * i) % V4_PX;
              if (x == st.d) {
                unpack8<T16>(__ldg(EL + off[i] + ro), r);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] -= r[e];
              }
              if (x == W - 1) {
                unpack8<T16>(__ldg(ER + off[i] + ro - st.d), r);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] -= r[e];
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], 0.f);
              v = pack8<T16>(a);
            } else {                                                          // relu(PL + PR); dead items stay zero
              v = make_uint4(add_relu2<T16>(vl[i].x, vr[i].x), add_relu2<T16>(vl[i].y, vr[i].y),
                             add_relu2<T16>(vl[i].z, vr[i].z), add_relu2<T16>(vl[i].w, vr[i].w));
            }
            dst[128 * i] = v;                                                 // ring row = [chunk][pixel] = linear in item
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy (UMMA)
          mbar_arrive(full + 8 * slot);
          if (rec) { c_wait += c1 - c0; c_load += clock64() - c1; }
        }
      }
      if (rec) {
        atomicAdd(prof + 3 + 3 * grp, (unsigned long long)c_wait);
        atomicAdd(prof + 4 + 3 * grp, (unsigned long long)c_load);
        atomicAdd(prof + 5 + 3 * grp, (unsigned long long)(clock64() - c_beg));
      }
    } else if (warp == 5 && lane == 0) {
      uint32_t rid0 = 0;
      long long c_wait = 0;
      const long long c_beg = prof ? clock64() : 0;
      for (int s = blockIdx.x; s < g.strips; s += gridDim.x, rid0 += (uint32_t)H) {
        const V4Strip st = v4_strip(s, g);
        for (int y = 0; y < H; ++y) {
          const uint32_t rid = rid0 + (uint32_t)y, slot = rid % V4_NROWS;
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(empty + 8 * slot, ((rid / V4_NROWS) & 1) ^ 1);
          if (prof) c_wait += clock64() - c0;
          mbar_expect_tx(full + 8 * slot, V4_ROW_BYTES);
          tma_load_5d(smem_u32(ring + slot * V4_ROW_BYTES), &tmIn, full + 8 * slot, 0, st.x0 - 1, y, 0, st.d * g.B + st.b);
        }
      }
      if (prof) {
        atomicAdd(prof + 3, (unsigned long long)c_wait);
        atomicAdd(prof + 5, (unsigned long long)(clock64() - c_beg));
      }
    }
  } else {
    // ==================================================================================== epilogue
    uint32_t it = 0;
    const bool rec = prof && threadIdx.x == 0;
    long long c_wait = 0;
    const long long c_beg = rec ? clock64() : 0;
    for (int s = blockIdx.x; s < g.strips; s += gridDim.x) {
      const V4Strip st = v4_strip(s, g);
      const int x = st.x0 + 32 * warp + lane;
      const bool keep = x >= st.d;
      for (int y = 0; y < H; ++y, ++it) {
        const uint32_t ab = it % V4_NACC;
        const long long c0 = rec ? clock64() : 0;
        mbar_wait(accfull + 8 * ab, (it / V4_NACC) & 1);
        if (rec) c_wait += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ab * NOUT + ((uint32_t)(32 * warp) << 16);
        uint32_t r[NOUT / 16][16];
#pragma unroll
        for (int u = 0; u < NOUT / 16; ++u) tmem_ld16(taddr + 16 * u, r[u]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(accempty + 8 * ab);
        if (x >= W) continue;
        if constexpr (GEN) {
          // layer 2: + folded bias, ReLU, x < d -> 0, 16-bit, [d][b][chunk = 4j + co/8][y][x][8]
          uint4* o = act_out + ((((int64_t)st.d * g.B + st.b) * 8 + 4 * st.j) * H + y) * (int64_t)W + x;
#pragma unroll
          for (int cc = 0; cc < NOUT / 8; ++cc) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int co = 8 * cc + e;
              v[e] = keep ? fmaxf(__uint_as_float(r[co >> 4][co & 15]) + fsm[co], 0.f) : 0.f;
            }
            o[(int64_t)cc * H * W] = pack8<T16>(v);
          }
        } else {
          // layer 3: + folded bias, ReLU, then volume11 (1x1 conv 16 -> 1, BN folded) + ReLU, x < d -> 0
          float acc = fsm[48];
#pragma unroll
          for (int co = 0; co < NOUT; ++co)
            acc = fmaf(fsm[32 + co], fmaxf(__uint_as_float(r[co >> 4][co & 15]) + fsm[co], 0.f), acc);
          vol[(((int64_t)st.b * g.D + st.d) * H + y) * (int64_t)W + x] = from_f<Tout>(keep ? fmaxf(acc, 0.f) : 0.f);
        }
      }
    }
    if (rec) {
      atomicAdd(prof + 9, (unsigned long long)c_wait);
      atomicAdd(prof + 10, (unsigned long long)(clock64() - c_beg));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int NOUT>
constexpr size_t v4_smem_bytes() {
  return 1024 + (size_t)V4_NROWS * V4_ROW_BYTES + 9 * V4_KCH * NOUT * 16 + 64 * 4 + (2 * V4_NROWS + 2 * V4_NACC) * 8 + 16;
}

struct V4Workspace {       // byte offsets into the caller's workspace
  size_t pl, pr, el, er, act, total;
};
static V4Workspace v4_workspace(int64_t B, int64_t H, int64_t W, int64_t D) {
  V4Workspace w;
  const size_t map = (size_t)B * 16 * H * W * 16;          // [b][16 chunks][H][W] x 16 B
  const size_t act = (size_t)D * B * 8 * H * W * 16;       // [d][b][8 chunks][H][W] x 16 B
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  w.pl = 0; w.pr = up(map); w.el = w.pr + up(map); w.er = w.el + up(map); w.act = w.er + up(map);
  w.total = w.act + up(act);
  return w;
}

}  // namespace rsm

using namespace rsm;

extern "C" int64_t rsm_v4_volume_workspace(int64_t B, int64_t H, int64_t W, int64_t D) {
  if (B < 0 || H < 0 || W < 0 || D < 0) return -1;
  return (int64_t)v4_workspace(B, H, W, D).total;
}

template <typename Tin, typename T16>
static int v4_run(const rsm_feat& left, const rsm_feat& right, const rsm_v4_weights& w, void* out, void* workspace,
                  int64_t B, int64_t H, int64_t W, int64_t D, int fmt, cudaStream_t st, unsigned long long* prof) {
  const V4Workspace ws = v4_workspace(B, H, W, D);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  uint4 *PL = (uint4*)(base + ws.pl), *PR = (uint4*)(base + ws.pr), *EL = (uint4*)(base + ws.el), *ER = (uint4*)(base + ws.er),
        *ACT = (uint4*)(base + ws.act);
  // K1
  {
    const dim3 grid((unsigned)ceil_div(W, 32), (unsigned)H, (unsigned)(2 * B));
    v4_premap_kernel<Tin, T16><<<grid, 256, 0, st>>>(view_of(left), view_of(right), w.w1, w.t1, PL, PR, EL, ER, (int)H, (int)W);
    if (int rc = finish_launch("rsm_v4_volume_fwd(premap)")) return rc;
  }
  V4Geom g;
  g.B = (int)B; g.H = (int)H; g.W = (int)W; g.D = (int)D; g.xtiles = (int)ceil_div(W, V4_TM); g.fmt = fmt;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  // K2
  {
    g.nj = 2; g.strips = (int)(B * D * 2 * g.xtiles);
    auto k = v4_conv_kernel<V4_C2, true, T16, Tin>;
    constexpr size_t smem = v4_smem_bytes<V4_C2>();
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return finish_launch("rsm_v4_volume_fwd(attr2)");
    const unsigned grid = (unsigned)(g.strips < kNumSMs ? g.strips : kNumSMs);
    k<<<grid, 13 * 32, smem, st>>>(PL, PR, EL, ER, tm, (const uint4*)w.w2, w.t2, nullptr, nullptr, ACT, (Tin*)nullptr, g, prof);
    if (int rc = finish_launch("rsm_v4_volume_fwd(conv2)")) return rc;
  }
  // K3
  {
    const TmapEncodeFn enc = tmap_encoder();
    if (!enc) return RSM_ERR_UNSUPPORTED_CONFIG;
    const cuuint64_t gdim[5] = {8, (cuuint64_t)W, (cuuint64_t)H, 8, (cuuint64_t)(D * B)};
    const cuuint64_t gstr[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 16 * 8};
    const cuuint32_t box[5] = {8, (cuuint32_t)V4_PX, 1, 8, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&tm, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ACT, gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RSM_ERR_UNSUPPORTED_CONFIG;
    g.nj = 1; g.strips = (int)(B * D * g.xtiles);
    auto k = v4_conv_kernel<V4_C3, false, T16, Tin>;
    constexpr size_t smem = v4_smem_bytes<V4_C3>();
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return finish_launch("rsm_v4_volume_fwd(attr3)");
    const unsigned grid = (unsigned)(g.strips < kNumSMs ? g.strips : kNumSMs);
    k<<<grid, 6 * 32, smem, st>>>(nullptr, nullptr, nullptr, nullptr, tm, (const uint4*)w.w3, w.t3, w.w11, w.t11, nullptr, (Tin*)out, g, prof ? prof + 16 : nullptr);
    return finish_launch("rsm_v4_volume_fwd(conv3)");
  }
}

static int v4_entry(rsm_feat left, rsm_feat right, rsm_v4_weights w, void* out, void* workspace, int64_t B,
                    int64_t C, int64_t H, int64_t W, int64_t D, int in_dtype, int op_dtype, int device,
                    void* stream, unsigned long long* prof) {
  if (B < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  if (C != V4_C) return RSM_ERR_UNSUPPORTED_CONFIG;                 // the module's Conv3d depths (8, 4, 2) need 2C = 64
  if (op_dtype != RSM_F16 && op_dtype != RSM_BF16) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (!valid_dtype(in_dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (B * H * W * D == 0) return RSM_OK;
  if (B > 32767 || H > 65535 || B * D * 2 * ceil_div(W, V4_TM) > 2147483647LL || B * 16 * H * W > 2147483647LL)
    return RSM_ERR_INVALID_SHAPE;
  if (!left.data || !right.data || !out || !workspace || !w.w1 || !w.t1 || !w.w2 || !w.t2 || !w.w3 || !w.t3 || !w.w11 || !w.t11)
    return RSM_ERR_NULL_POINTER;
  if (!aligned_to(workspace, 256) || !aligned_to(w.w2, 16) || !aligned_to(w.w3, 16)) return RSM_ERR_MISALIGNED;
  DeviceGuard guard(device);
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int fmt = op_dtype == RSM_F16 ? 0 : 1;
  if (in_dtype == RSM_F32) {
    return fmt == 0 ? v4_run<float, __half>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof)
                    : v4_run<float, __nv_bfloat16>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof);
  }
  if (in_dtype == RSM_F16) {
    return fmt == 0 ? v4_run<__half, __half>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof)
                    : v4_run<__half, __nv_bfloat16>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof);
  }
  return fmt == 0 ? v4_run<__nv_bfloat16, __half>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof)
                  : v4_run<__nv_bfloat16, __nv_bfloat16>(left, right, w, out, workspace, B, H, W, D, fmt, st, prof);
}

extern "C" int rsm_v4_volume_fwd(rsm_feat left, rsm_feat right, rsm_v4_weights w, void* out, void* workspace, int64_t B,
                                 int64_t C, int64_t H, int64_t W, int64_t D, int in_dtype, int op_dtype, int device,
                                 void* stream) {
  return v4_entry(left, right, w, out, workspace, B, C, H, W, D, in_dtype, op_dtype, device, stream, nullptr);
}

// diagnostic twin: `prof` = 32 zero-initialised uint64 on the device; the kernels add clock64 cycles per role
// ([0..10] layer 2, [16..26] layer 3: issuer {waiting for rows, waiting for an accumulator, total}, producers
// {waiting for a slot, working, total} x 2 groups, [8] rows issued, epilogue {waiting, total}), summed over CTAs
extern "C" int rsm_v4_volume_fwd_profile(rsm_feat left, rsm_feat right, rsm_v4_weights w, void* out, void* workspace,
                                         int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int in_dtype,
                                         int op_dtype, int device, void* stream, uint64_t* prof) {
  if (!prof) return RSM_ERR_NULL_POINTER;
  return v4_entry(left, right, w, out, workspace, B, C, H, W, D, in_dtype, op_dtype, device, stream,
                  reinterpret_cast<unsigned long long*>(prof));
}
