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

}  // namespace rsm
