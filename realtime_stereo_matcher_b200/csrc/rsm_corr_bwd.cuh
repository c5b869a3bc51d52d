// Tiled adjoint of the correlation volumes (inner product / mean correlation / group-wise), SIMT fp32.
//
//   gL[c,x]  = s * sum_d gV[g(c), d, x]      * R[c, x - d]          (SIDE_LEFT)
//   gR[c,x'] = s * sum_d gV[g(c), d, x' + d] * L[c, x' + d]         (SIDE_RIGHT)
//
// Same shape as the forward: a CTA owns a 64-pixel row segment of one (n, y) and a block of <= 32
// channels of ONE group; per chunk of <= 64 disparities it stages the gradient tile gV[d][x] and the
// other feature's window in shared memory as fp32, and every thread accumulates a 4(x) x 4(c)
// register tile, walking the disparities four at a time so that all shared-memory reads are aligned
// 128-bit loads (4-6 FMA per LDS.128).  Atomic-free and deterministic.
//
// Thread layout: 16 pixel quads x 8 "tc" slots.  With a full 32-channel block a tc slot is a channel
// quad.  Narrow blocks (group-wise volumes: 4 or 8 channels per group) would leave most slots idle, so
// the slots are re-used to split the DISPARITY quads of a chunk dsplit = 8 / #channel-quads ways; the
// partial register tiles are then reduced through shared memory in a fixed order (still deterministic).
#pragma once

namespace rsm {

constexpr int BW_TX = 64;    // pixels per CTA
constexpr int BW_CB = 32;    // channels per CTA (<= one group)
constexpr int BW_DCH = 64;   // disparities per staged chunk

enum { SIDE_LEFT = 0, SIDE_RIGHT = 1 };

// gradient element gV[d, x] of (n, group, y) in either output layout
template <typename Tout, int LAYOUT>
struct GradView {
  const Tout* base;
  int64_t sd, sx;
  __device__ __forceinline__ GradView(const Tout* gout, const CorrGeom& g, int64_t n, int grp, int y) {
    if constexpr (LAYOUT == LAYOUT_NDHW) {
      base = gout + ((int64_t)n * g.D * g.H + y) * g.W; sd = (int64_t)g.H * g.W; sx = 1;
    } else {
      base = gout + (((int64_t)n * g.G + grp) * g.H + y) * (int64_t)g.W * g.D; sd = 1; sx = g.D;
    }
  }
  __device__ __forceinline__ float at(int d, int x) const { return to_f(__ldg(base + d * sd + x * sx)); }
};

template <typename Tin, typename Tout, int LAYOUT, int SIDE>
__global__ void __launch_bounds__(16 * (BW_CB / 4))
corr_bwd_tiled_kernel(const Tout* __restrict__ gout, FeatView L, FeatView R, Tin* __restrict__ gdst, CorrGeom g,
                      int cblocks, int cb_size) {
  extern __shared__ __align__(16) float smem[];
  // SIDE_LEFT : sG[BW_DCH][TX]          gradient rows x0..x0+TX;      sF = R window [x0-dc0-DCH, x0-dc0+TX)
  // SIDE_RIGHT: sG[BW_DCH][TX + DCH]    gradient rows x0+dc0.. ;      sF = L window [x0+dc0, x0+dc0+TX+DCH)
  constexpr int GW = SIDE == SIDE_LEFT ? BW_TX : BW_TX + BW_DCH;
  constexpr int GP = GW + 4;   // row pitch: 16-byte aligned rows, and = 4 (mod 32) so that the transposing
                               // stage of the D-innermost layout (4 pixels x 8 disparities per warp) is conflict-free
  constexpr int FW = BW_TX + BW_DCH;
  float* sG = smem;
  float* sF = smem + BW_DCH * GP;

  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % g.xtiles); bid /= g.xtiles;
  const int cb = (int)(bid % cblocks); bid /= cblocks;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * BW_TX;
  const int c0 = cb * cb_size;                 // first channel of this block (never straddles a group)
  const int grp = c0 / g.cpg;
  const int ncb = min(cb_size, g.C - c0);
  const int tx = threadIdx.x & 15, tc = threadIdx.x >> 4;
  // channel quads in this block, and how many ways the disparity quads are split over the spare tc slots
  const int cq = (ncb + 3) >> 2;
  const int dsplit = cq == 1 ? 8 : cq == 2 ? 4 : 1;   // (a 2-way split of 4 quads measured slower than none)
  const int cquad = dsplit == 1 ? tc : tc % (8 / dsplit), dpart = dsplit == 1 ? 0 : tc / (8 / dsplit);
  const GradView<Tout, LAYOUT> gv(gout, g, n, grp, y);
  const FeatView& F = SIDE == SIDE_LEFT ? R : L;
  const Tin* __restrict__ pf = reinterpret_cast<const Tin*>(F.data) + n * F.sn + (int64_t)y * F.sh + (int64_t)c0 * F.sc;

  float acc[4][4];   // [channel j][pixel i]
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  for (int dc0 = 0; dc0 < g.D; dc0 += BW_DCH) {
    __syncthreads();
    // ---- stage the gradient tile (zero outside [0,D) x [0,W)) and the feature window (zero outside [0,W))
    const int gx0 = SIDE == SIDE_LEFT ? x0 : x0 + dc0;
    for (int e = threadIdx.x; e < BW_DCH * GW; e += blockDim.x) {
      int dl, xx;
      if constexpr (LAYOUT == LAYOUT_NDHW) { dl = e / GW; xx = e - dl * GW; }      // x fastest: coalesced rows
      else {   // d fastest in memory: a warp takes 4 pixels x 8 consecutive disparities (32-byte runs)
        static_assert(BW_DCH == 64, "index split below assumes 64 disparities per chunk");
        dl = (e >> 2) & 63; xx = (e & 3) | ((e >> 8) << 2);
      }
      const int d = dc0 + dl, x = gx0 + xx;
      const bool valid = d < g.D && x < g.W;
      if constexpr (sizeof(Tout) == 4) {   // fp32 gradient: LDGSTS with zero-fill, all loads of the tile in flight
        const Tout* src = valid ? gv.base + d * gv.sd + x * gv.sx : gv.base;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sG + dl * GP + xx)),
                     "l"(src), "r"(valid ? 4 : 0)
                     : "memory");
      } else {
        sG[dl * GP + xx] = valid ? gv.at(d, x) : 0.f;
      }
    }
    const int fx0 = SIDE == SIDE_LEFT ? x0 - dc0 - BW_DCH : x0 + dc0;
    for (int e = threadIdx.x; e < ncb * FW; e += blockDim.x) {
      const int c = e / FW, j = e - c * FW;
      const int x = fx0 + j;
      const bool valid = x >= 0 && x < g.W;
      if constexpr (sizeof(Tin) == 4) {
        const Tin* src = valid ? pf + (int64_t)c * F.sc + (int64_t)x * F.sw : pf;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sF + c * FW + j)),
                     "l"(src), "r"(valid ? 4 : 0)
                     : "memory");
      } else {
        sF[c * FW + j] = valid ? to_f(__ldg(pf + (int64_t)c * F.sc + (int64_t)x * F.sw)) : 0.f;
      }
    }
    if constexpr (sizeof(Tin) == 4 || sizeof(Tout) == 4) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (4 * cquad >= ncb) continue;
    // ---- accumulate: disparities four at a time (d = dc0 + 4q + r)
    for (int q = dpart; q < BW_DCH / 4; q += dsplit) {
      if (dc0 + 4 * q >= g.D) break;
      if constexpr (SIDE == SIDE_LEFT) {
        float gq[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 t = *reinterpret_cast<const float4*>(sG + (4 * q + r) * GP + 4 * tx);
          gq[r][0] = t.x; gq[r][1] = t.y; gq[r][2] = t.z; gq[r][3] = t.w;
        }
        // R[x - d]: window index = DCH + 4tx + i - 4q - r = (DCH + 4tx - 4q - 4) + (4 + i - r)
        const int wb = BW_DCH + 4 * tx - 4 * q - 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4* wp = reinterpret_cast<const float4*>(sF + (4 * cquad + j) * FW + wb);
          const float4 w0 = wp[0], w1 = wp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(gq[r][i], w[4 + i - r], acc[j][i]);
        }
      } else {
        // u = x' + d: window index = 4tx + 4q + (i + r) in both sG (row d) and sF
        const int wb = 4 * tx + 4 * q;
        float p[4][8];   // p[r][k] = gV[d = 4q + r][wb + k]
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4* gp = reinterpret_cast<const float4*>(sG + (4 * q + r) * GP + wb);
          const float4 a = gp[0], b = gp[1];
          p[r][0] = a.x; p[r][1] = a.y; p[r][2] = a.z; p[r][3] = a.w;
          p[r][4] = b.x; p[r][5] = b.y; p[r][6] = b.z; p[r][7] = b.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4* lp = reinterpret_cast<const float4*>(sF + (4 * cquad + j) * FW + wb);
          const float4 l0 = lp[0], l1 = lp[1];
          const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(p[r][i + r], l[i + r], acc[j][i]);
        }
      }
    }
  }
  // ---- scale and store (x contiguous)
  const float cnt = g.mean ? (float)g.cpg : 1.f;
  const int xb = x0 + 4 * tx;
  if (dsplit == 1) {
    if (xb >= g.W) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 4 * tc + j;
      if (4 * tc + j >= ncb) break;
      Tin* o = gdst + (((int64_t)n * g.C + c) * g.H + y) * g.W + xb;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (xb + i < g.W) o[i] = from_f<Tin>(acc[j][i] / cnt);
    }
    return;
  }
  // ---- disparity-split blocks: part[tc][v = 4j + i][tx], reduced over the dsplit slots of a channel quad
  // in ascending slot order; slot dpart of (cquad, tx) finishes values v0 .. v0 + 16/dsplit - 1
  __syncthreads();
  float* part = smem;   // 8 * 16 * 16 floats = 8 KB <= the gradient tile it overlays
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) part[(tc * 16 + 4 * j + i) * 16 + tx] = acc[j][i];
  __syncthreads();
  if (4 * cquad >= ncb) return;
  const int ncq = 8 / dsplit, nv = 16 / dsplit;
  for (int v = dpart * nv; v < (dpart + 1) * nv; ++v) {
    float sum = 0.f;
    for (int dp = 0; dp < dsplit; ++dp) sum += part[((dp * ncq + cquad) * 16 + v) * 16 + tx];
    const int j = v >> 2, i = v & 3;
    if (4 * cquad + j < ncb && xb + i < g.W)
      gdst[(((int64_t)n * g.C + c0 + 4 * cquad + j) * g.H + y) * g.W + xb + i] = from_f<Tin>(sum / cnt);
  }
}

// ================================================== inner product (N,D,H,W): 8(x) x BB_CPT(c) tiles
// Same contraction as corr_bwd_tiled_kernel with more arithmetic per shared-memory byte: CTA = 128
// pixels x 32 channels of one (n, y), thread tile 8(x) x BB_CPT(c), disparities eight at a time:
//     16 LDS.128 for the gradient block g[8 d][8 x] + 4 LDS.128 per channel for its 16-wide feature window
//     -> 8 x 8 tile: 512 FMA per 48 LDS.128 (10.7 per load); 8 x 4 tile: 256 per 32 (8); the 4x4 tile gets 5.3.
// For the right gradient the gradient tile is staged SKEWED, sG[d][x'] = gV[d][x' + d], which turns
// gR[c,x'] = sum_d gV[d][x'+d] L[c][x'+d] into the same Toeplitz form as the left one (window ascending
// instead of descending).  Rows are XOR-swizzled by 16-byte chunk (chunk ^= (chunk >> 3) & 1): lanes read
// 32-byte segments 32 bytes apart.  Staging: 16-byte LDGSTS with zero-fill wherever the source is
// chunk-aligned (everything except the fp32 skewed tile, which moves 4 bytes at a time).  16-bit tensors are
// widened on the way in: 16-byte loads (four in flight per thread) for the aligned tiles; the skewed tile a row
// per warp, eight rows in flight -- each lane takes the two aligned 8-byte quads around its four columns and
// funnel-shifts them by the row's misalignment (ncu: 750 -> 477 us for the right gradient at cfg2 C=64 against
// element-wise moves, which were latency-bound with one or two loads in flight).
// DCH = disparities staged per pass, a template parameter: 48 keeps the tiles at 47 KB (four CTAs per SM instead of
// three at 64 -- the kernel is latency-bound at 6 warps per SM, ncu: 42 % issue slots) and fits D = 48 / 96 / 192
// without dead rows; 64 where that wastes less.
// 128 threads = 16 pixel octets x 8 channel quads: the 8(x) x 4(c) thread tile has a worse FMA : LDS ratio than 8 x 8,
// but the kernel is latency-bound (ncu: 6 warps per SM, 42 % of the issue slots with 64-thread CTAs) and twice the
// warps won (cfg2 C=64 fp32 791 -> 699 us).
constexpr int BB_TX = 128, BB_CB = 32, BB_THREADS = 128, BB_CPT = BB_CB / (BB_THREADS / 16);   // channels per thread
constexpr size_t bb_smem_bytes(int dch) { return (size_t)(dch * BB_TX + BB_CB * (BB_TX + dch)) * sizeof(float); }

__device__ __forceinline__ int bb_swz(int chunk) { return chunk ^ ((chunk >> 3) & 1); }

// 8 consecutive 16-bit elements (one 16-byte load) widened to fp32 and stored as two swizzled 16-byte chunks
template <typename T> __device__ __forceinline__ float bits_to_f(uint32_t lo16);
template <> __device__ __forceinline__ float bits_to_f<__half>(uint32_t lo16) { return __half2float(__ushort_as_half((unsigned short)lo16)); }
template <> __device__ __forceinline__ float bits_to_f<__nv_bfloat16>(uint32_t lo16) { return __uint_as_float(lo16 << 16); }
template <> __device__ __forceinline__ float bits_to_f<float>(uint32_t) { return 0.f; }   // never used: fp32 moves by LDGSTS

template <typename T>
__device__ __forceinline__ void bb_store8(float* row, int ch2, const Vec16<T>& v, bool valid) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = valid ? to_f(v.v[i]) : 0.f;
  *reinterpret_cast<float4*>(row + 4 * bb_swz(2 * ch2)) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(row + 4 * bb_swz(2 * ch2 + 1)) = make_float4(f[4], f[5], f[6], f[7]);
}

// T = element type of the features, the gradient of the volume and the feature gradients (all equal here);
// fp32 moves by LDGSTS, 16-bit tensors by 16-byte loads widened on the way into shared memory.
template <typename T, int SIDE, int DCH>
__global__ void __launch_bounds__(BB_THREADS, 4)
inner_bwd_big_kernel(const T* __restrict__ gout, FeatView L, FeatView R, T* __restrict__ gdst, CorrGeom g,
                     int xtiles, int cblocks) {
  constexpr int FW = BB_TX + DCH;         // feature window: the tile plus DCH pixels of disparity reach
  extern __shared__ __align__(16) float smem[];
  float* sG = smem;                       // [DCH][BB_TX]
  float* sF = smem + DCH * BB_TX;      // [BB_CB][FW]
  int64_t bid = blockIdx.x;
  const int xt = (int)(bid % xtiles); bid /= xtiles;
  const int cb = (int)(bid % cblocks); bid /= cblocks;
  const int y = (int)(bid % g.H);
  const int64_t n = bid / g.H;
  const int x0 = xt * BB_TX, c0 = cb * BB_CB;
  const int ncb = min(BB_CB, g.C - c0);
  const int tx = threadIdx.x & 15, tc = threadIdx.x >> 4;
  const T* __restrict__ gbase = gout + ((int64_t)n * g.D * g.H + y) * g.W;         // + d * H * W + x
  const int64_t gsd = (int64_t)g.H * g.W;
  const FeatView& F = SIDE == SIDE_LEFT ? R : L;
  const T* __restrict__ pf = reinterpret_cast<const T*>(F.data) + n * F.sn + (int64_t)y * F.sh + (int64_t)c0 * F.sc;

  float acc[BB_CPT][8];   // [channel j][pixel i]
#pragma unroll
  for (int j = 0; j < BB_CPT; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;

  for (int dc0 = 0; dc0 < g.D; dc0 += DCH) {
    __syncthreads();
    // ---- gradient tile (the sum below reads whole groups of eight rows)
    const int nrows = min(DCH, (g.D - dc0 + 7) & ~7);
    if constexpr (SIDE == SIDE_LEFT && sizeof(T) == 4) {
      for (int e = threadIdx.x; e < DCH * (BB_TX / 4); e += BB_THREADS) {
        const int dl = e >> 5, ch = e & 31;
        const int d = dc0 + dl, x = x0 + 4 * ch;
        const bool valid = d < g.D && x < g.W;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sG + dl * BB_TX + 4 * bb_swz(ch))),
                     "l"(valid ? gbase + d * gsd + x : gbase), "r"(valid ? 16 : 0)
                     : "memory");
      }
    } else if constexpr (SIDE == SIDE_LEFT) {
      // W % 8 == 0: whole octets; four 16-byte loads in flight per thread before anything is widened
      for (int e0 = threadIdx.x; e0 < nrows * (BB_TX / 8); e0 += 4 * BB_THREADS) {
        Vec16<T> v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * BB_THREADS;
          const int d = dc0 + (e >> 4), x = x0 + 8 * (e & 15);
          ok[u] = d < g.D && x < g.W;
          if (ok[u]) v[u] = ldg16(gbase + d * gsd + x);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * BB_THREADS;
          if ((e >> 4) < nrows) bb_store8<T>(sG + (e >> 4) * BB_TX, e & 15, v[u], ok[u]);
        }
      }
    } else {
      if constexpr (sizeof(T) == 4) {
        for (int e = threadIdx.x; e < DCH * BB_TX; e += BB_THREADS) {
          const int dl = e >> 7, xx = e & 127;
          const int d = dc0 + dl, x = x0 + xx + d;             // skew: column x' holds gV[d][x' + d]
          const bool valid = d < g.D && x < g.W;
          float* dst = sG + dl * BB_TX + 4 * bb_swz(xx >> 2) + (xx & 3);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                       "l"(valid ? gbase + d * gsd + x : gbase), "r"(valid ? 4 : 0)
                       : "memory");
        }
      } else {
        // 16-bit: a warp moves one row at a time.  Lane l needs gV[d][x0 + d + 4l .. + 3]: the two aligned
        // quads around it, funnel-shifted by the row's misalignment (warp-uniform), widened, one 16-byte store
        // (eight rows = sixteen 8-byte loads in flight per lane: the staging is synchronous and latency-bound)
        const int lane = threadIdx.x & 31;
        constexpr int NW = BB_THREADS / 32, RB = 8;
        for (int dl0 = threadIdx.x >> 5; dl0 < nrows; dl0 += NW * RB) {
          uint2 lo[RB], hi[RB];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int d = dc0 + dl0 + NW * u;
            const int xa = ((x0 + d) & ~3) + 4 * lane;
            lo[u] = make_uint2(0u, 0u);
            hi[u] = make_uint2(0u, 0u);
            if (d < g.D) {
              const T* src = gbase + d * gsd;
              if (xa < g.W) lo[u] = __ldg(reinterpret_cast<const uint2*>(src + xa));          // W % 4 == 0: whole quads
              if (xa + 4 < g.W) hi[u] = __ldg(reinterpret_cast<const uint2*>(src + xa + 4));
            }
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int dl = dl0 + NW * u;
            if (dl >= nrows) break;
            const int s = (x0 + dc0 + dl) & 3;
            const bool wo = (s >> 1) != 0;
            const uint32_t w0 = wo ? lo[u].y : lo[u].x, w1 = wo ? hi[u].x : lo[u].y, w2 = wo ? hi[u].y : hi[u].x;
            const int bs = 16 * (s & 1);
            const uint32_t o0 = __funnelshift_r(w0, w1, bs), o1 = __funnelshift_r(w1, w2, bs);
            *reinterpret_cast<float4*>(sG + dl * BB_TX + 4 * bb_swz(lane)) =
                make_float4(bits_to_f<T>(o0 & 0xffffu), bits_to_f<T>(o0 >> 16), bits_to_f<T>(o1 & 0xffffu), bits_to_f<T>(o1 >> 16));
          }
        }
      }
    }
    // ---- feature windows: left side R[c][x0 - dc0 - 64 + j], right side L[c][x0 + dc0 + j]
    const int fx0 = SIDE == SIDE_LEFT ? x0 - dc0 - DCH : x0 + dc0;
    if constexpr (sizeof(T) == 4) {
      for (int e = threadIdx.x; e < (FW / 4); e += BB_THREADS) {
        const int x = fx0 + 4 * e;
        const bool valid = x >= 0 && x < g.W;
        const T* src = valid ? pf + x : pf;
        const int64_t step = valid ? F.sc : 0;
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(sF + 4 * bb_swz(e));
        for (int c = 0; c < ncb; ++c)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + (uint32_t)(c * FW * 4)), "l"(src + c * step),
                       "r"(valid ? 16 : 0)
                       : "memory");
      }
    } else {
      // fx0 % 8 == 0 and W % 8 == 0: whole octets, four loads in flight (deeper batches measured slower)
      const int total = ncb * (FW / 8);
      for (int e0 = threadIdx.x; e0 < total; e0 += 4 * BB_THREADS) {
        Vec16<T> v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * BB_THREADS;
          const int c = e / (FW / 8), x = fx0 + 8 * (e - c * (FW / 8));
          ok[u] = e < total && x >= 0 && x < g.W;
          if (ok[u]) v[u] = ldg16(pf + (int64_t)c * F.sc + x);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * BB_THREADS;
          const int c = e / (FW / 8);
          if (e < total) bb_store8<T>(sF + c * FW, e - c * (FW / 8), v[u], ok[u]);
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (BB_CPT * tc >= ncb) continue;
    for (int d0 = 0; d0 < DCH; d0 += 8) {
      if (dc0 + d0 >= g.D) break;
      float gq[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float* row = sG + (d0 + r) * BB_TX;
        const float4 a = *reinterpret_cast<const float4*>(row + 4 * bb_swz(2 * tx));
        const float4 b = *reinterpret_cast<const float4*>(row + 4 * bb_swz(2 * tx + 1));
        gq[r][0] = a.x; gq[r][1] = a.y; gq[r][2] = a.z; gq[r][3] = a.w;
        gq[r][4] = b.x; gq[r][5] = b.y; gq[r][6] = b.z; gq[r][7] = b.w;
      }
      // window chunk: left  w[k] = R[x + i - d], k = 8 + i - r  (window starts at 64 + 8 tx - d0 - 8)
      //               right w[k] = L[x' + i + d], k = i + r      (window starts at 8 tx + d0)
      const int wch = SIDE == SIDE_LEFT ? (DCH + 8 * tx - d0 - 8) >> 2 : (8 * tx + d0) >> 2;
#pragma unroll
      for (int j = 0; j < BB_CPT; ++j) {
        if (BB_CPT * tc + j >= ncb) break;
        const float* frow = sF + (BB_CPT * tc + j) * FW;
        float w[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 t = *reinterpret_cast<const float4*>(frow + 4 * bb_swz(wch + k));
          w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            acc[j][i] = fmaf(gq[r][i], SIDE == SIDE_LEFT ? w[8 + i - r] : w[i + r], acc[j][i]);
      }
    }
  }
  // ---- scale and store (x contiguous)
  const float cnt = g.mean ? (float)g.cpg : 1.f;
  const int xb = x0 + 8 * tx;
  if (xb >= g.W) return;
#pragma unroll
  for (int j = 0; j < BB_CPT; ++j) {
    if (BB_CPT * tc + j >= ncb) break;
    T* o = gdst + (((int64_t)n * g.C + c0 + BB_CPT * tc + j) * g.H + y) * g.W + xb;
    if constexpr (sizeof(T) == 4) {
      __stcs(reinterpret_cast<float4*>(o), make_float4(acc[j][0] / cnt, acc[j][1] / cnt, acc[j][2] / cnt, acc[j][3] / cnt));
      if (xb + 4 < g.W)
        __stcs(reinterpret_cast<float4*>(o + 4), make_float4(acc[j][4] / cnt, acc[j][5] / cnt, acc[j][6] / cnt, acc[j][7] / cnt));
    } else {
      Vec16<T> v;                                             // W % 8 == 0: the whole octet is inside the row
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = from_f<T>(acc[j][i] / cnt);
      stcs16(o, v);
    }
  }
}

}  // namespace rsm
