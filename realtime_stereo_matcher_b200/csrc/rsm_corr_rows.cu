// Fused inner-product -> disparity regression on tcgen05, row-streaming form (16-bit features, C <= 128, D <= 384):
// soft-argmax / argmin / argmax / log-sum-exp of the correlation volume of make_correlation_volume
// (model/mobile_disp_net_c.py:188-205) / TorchInnerProductCost (cost_volume/inner_product.py:11-42) without the
// (N,D,H,W) volume ever existing, every feature element leaving L2 ONCE.
//
// Per epipolar row the volume is the band 0 <= x - x' < D of P[x, x'] = sum_c L[c, x] R[c, x'].  The first fused kernel
// (inner_tc_kernel<EPI_REGRESS>, rsm_corr_tc.cu) staged, per 128-pixel tile and disparity chunk, the left tile plus a
// (128 + chunk)-pixel right window: 2.4-3.2x the algorithmic bytes through the L2 -> shared-memory path, and its
// epilogue un-skewed the band through shared memory.  Here:
//   * operands are 64-pixel ATOMS = one TMA box of 64 pixels x C channels, SWIZZLE_128B (the MN-major UMMA atom:
//     128-byte rows, 8-channel groups 1 KB apart).  The right row streams through a ring of atoms: the window of
//     x-tile i is atoms [2i, 2i + na) of the row stream (na = 2 + ceil(D/64)), the next tile re-uses all but two of
//     them in place -- the UMMA B descriptor simply points at ring slots.  The left tile (two atoms) has its own ring.
//   * a window is cut into one or two GROUPS of <= 4 atoms; a group is one accumulator of <= 256 columns in one half
//     of TMEM (two halves: the MMAs of group g+1 overlap the epilogue of group g) and one tcgen05.mma per 16 channels
//     and run of contiguous ring slots (N = 64 x atoms, almost always the whole group).  MMA work per tile is
//     128 x (128 + Dp) x C instead of 128 x (128 + chunk) x C per disparity chunk: 0.63x at D = 192.
//   * the epilogue never leaves the TMEM lane: lane r (pixel x0 + r) holds P[x0 + r, x0 - Dp + j] in column j, i.e.
//     disparity d = r + Dp - j.  Each warp scans its lanes' own columns 8 at a time with the lean online softmax;
//     columns outside [0, D) for a lane are masked (one compare + select per value), which only happens in the two
//     31-column edges of a quadrant's range.  j ascends = d descends, so ties take the LATER column (torch: first
//     index wins).
// Warp roles (352 threads, one persistent CTA per SM):
//   warps 0-7   epilogue: TMEM lane quadrant q = warp % 4 (also the warp's scheduler), the 8-column chunks of the
//               quadrant's range dealt alternately to its two warps; partial states merged through shared memory
//   warp 8      UMMA issuer (whole warp converged, one elected lane issues)
//   warps 9, 10 TMA producers (one lane each): right-atom ring, left-tile ring
// mbarriers: a_full/a_empty (left tile ring), b_full/b_empty (right atom ring), t_full/t_empty (TMEM halves).  All
// waits are bounded and trap on expiry (rsm_tc.cuh).
// Measured while building it (clock64 stamps per role and atom, first version: one hand-off per 64-column atom, 12
// epilogue warps): a warp that only waits, fences and arrives needs 300-450 cycles per hand-off, a single-lane
// producer with divisions in its loop ~600 cycles per box, and one warp alone needs ~500 cycles per 8-column chunk
// (TMEM round trip + a serial scan) -- so hand-offs are per group, ring positions are kept incrementally, and every
// warp scans two chunks at a time into two independent states.
// Where the time goes now (experiment builds, config 2 C = 64 / config 4 C = 128 D = 192 x 8 images, same box): loads +
// hand-offs alone, no MMAs and no scan, 71 / 110 us = 4.0 / 4.8 TB/s -- what TMA boxes of 128-byte rows gathered from C
// channel planes deliver from HBM, whatever the ring depths (2 .. 8 left tiles, 5 .. 16 right atoms: no change) and
// whether one lane or two issue them; + MMAs 77 / 131 us; + scan 98 / 188 us.  Taking the exponentials or the TMEM
// loads OUT of the scan changes nothing (<= 5 %): the scan's cost is its ~75 plain instructions per 16 values.
#include <cuda.h>

#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int RR_TM = 128;          // left pixels per x-tile (UMMA M)
constexpr int RR_ATOM = 64;         // pixels per operand atom
constexpr int RR_GATOMS = 4;        // atoms per accumulator group (256 TMEM columns = one half)
constexpr int RR_MAXA = 4;          // left-tile ring slots (upper bound)
constexpr int RR_MAXB = 16;         // right-atom ring slots (upper bound)
constexpr int RR_NSPLIT = 2;        // epilogue warps per TMEM lane quadrant (3 and 4 measured the same or slower)
constexpr int RR_EPI_WARPS = 4 * RR_NSPLIT;
constexpr int RR_THREADS = 32 * (RR_EPI_WARPS + 3);
constexpr int RR_PART_BYTES = 2 * RR_NSPLIT * 128 * 8 * 4;   // parked partial states: 2 buffers x 2 parts x 128 lanes x 8 words
constexpr int RR_BAR_BYTES = 512;   // 2*4 + 2*16 + 2*2 = 44 mbarriers + the TMEM address slot
constexpr int RR_SMEM_MAX = 227 * 1024;

struct RrGeom {
  int C, H, W, D;
  int Dp;          // D rounded up to whole atoms
  int na;          // atoms per x-tile window = 2 + Dp / 64 (<= 8)
  int ng;          // accumulator groups per window (1 or 2)
  int gsz0;        // atoms in group 0 (group 1 has na - gsz0)
  int nabuf, nb;   // ring slots in use
  int xtiles;
  int fmt;         // 0 = fp16, 1 = bf16
  int mean, pow2;
  int atom_bytes;  // C * 128
  int64_t rows, tiles;
};

struct RrOut {
  float* soft;
  int64_t* amin;
  int64_t* amax;
  float* lse;
};

// The x-tiles of one CTA, in order: a contiguous range of (n, y, xt) tiles.  (Dealing whole rows round-robin over the
// CTAs, so that the grid reads neighbouring rows of every channel plane at any time, measured the same.)
struct RowTile {
  int n, y, xt;
  __device__ __forceinline__ void advance(const RrGeom& g) {
    if (++xt < g.xtiles) return;
    xt = 0;
    if (++y < g.H) return;
    y = 0; ++n;
  }
};
struct TileRange {
  RowTile first;
  uint32_t ntl;
};
__device__ __forceinline__ TileRange tile_range(const RrGeom& g) {
  TileRange tr;
  const int64_t per = (g.tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_beg = min((int64_t)blockIdx.x * per, g.tiles), t_end = min(t_beg + per, g.tiles);
  const uint32_t t = (uint32_t)t_beg, row = t / (uint32_t)g.xtiles;
  tr.ntl = (uint32_t)(t_end - t_beg);
  tr.first.xt = (int)(t - row * (uint32_t)g.xtiles);
  tr.first.n = (int)(row / (uint32_t)g.H);
  tr.first.y = (int)(row - (uint32_t)tr.first.n * (uint32_t)g.H);
  return tr;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// running per-lane state: softmax max (raw units) / sum / weighted sum, extrema with their disparities, first NaN
struct RrState {
  float m, s, ws, minv, maxv;
  int mini, maxi, nani;
  __device__ __forceinline__ void reset() {
    m = -1e30f; s = 0.f; ws = 0.f; minv = INFINITY; maxv = -INFINITY;
    mini = 0x7fffffff; maxi = 0x7fffffff; nani = 0x7fffffff;
  }
};

// 8 consecutive window columns of this lane: disparities dtop, dtop-1, ..., dtop-7.
// MODE 0: every column is a disparity in [0, D) for every lane of the warp (interior of the band);
// MODE 4: an edge of the quadrant's range: columns whose disparity falls outside [0, D) for this lane do not take part;
// MODE 3: as 4, and columns below jfill (x' < 0) read as the reference's fill value 0 (first tiles of a row only).
// k2 = log2(e) * scale folded into the exponent; extrema are tracked on the raw sums (the scale is a positive power of
// two there, or has been applied by an exact division -- DIV -- when it is not).
template <bool EXT, bool DIV, int MODE>
__device__ __forceinline__ void rr_chunk8(RrState& st, const uint32_t (&raw)[8], int c0, int dtop, int D, int jfill, float k2, float cnt) {
  float vx[8], vn[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float f = __uint_as_float(raw[k]);
    if (DIV) f = __fdiv_rn(f, cnt);
    bool ok = true;
    if (MODE == 3 && c0 + k < jfill) f = 0.f;                  // warp-uniform predicate
    if (MODE != 0) ok = (unsigned)(dtop - k) < (unsigned)D;
    vx[k] = (MODE == 0 || ok) ? f : -INFINITY;
    if (EXT) vn[k] = (MODE == 0 || ok) ? f : INFINITY;
  }
  const float cmax = fmaxf(fmaxf(fmaxf(vx[0], vx[1]), fmaxf(vx[2], vx[3])), fmaxf(fmaxf(vx[4], vx[5]), fmaxf(vx[6], vx[7])));
  const float mn = fmaxf(st.m, cmax), mnl = mn * k2;
  const float a = fast_exp2((st.m - mn) * k2);      // (not fma(m, k2, -mnl): m may still be the -1e30 start value)
  float e[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) e[k] = fast_exp2(fmaf(vx[k], k2, -mnl));
  const float S = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
  const float T = (fmaf(2.f, e[2], e[1]) + fmaf(3.f, e[3], 4.f * e[4])) + (fmaf(5.f, e[5], 6.f * e[6]) + 7.f * e[7]);   // sum k e_k
  st.s = fmaf(st.s, a, S);
  st.ws = fmaf(st.ws, a, fmaf((float)dtop, S, -T));                                      // sum (dtop - k) e_k
  st.m = mn;
  if (EXT) {
    const float cmin = fminf(fminf(fminf(vn[0], vn[1]), fminf(vn[2], vn[3])), fminf(fminf(vn[4], vn[5]), fminf(vn[6], vn[7])));
    int imin = 0, imax = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) {                     // LAST k attaining the extremum = smallest disparity
      imin = vn[k] == cmin ? k : imin;
      imax = vx[k] == cmax ? k : imax;
    }
    if (MODE != 0) {
      if (cmin <= st.minv && cmin < INFINITY) { st.minv = cmin; st.mini = dtop - imin; }
      if (cmax >= st.maxv && cmax > -INFINITY) { st.maxv = cmax; st.maxi = dtop - imax; }
    } else {
      if (cmin <= st.minv) { st.minv = cmin; st.mini = dtop - imin; }
      if (cmax >= st.maxv) { st.maxv = cmax; st.maxi = dtop - imax; }
    }
    if (S != S) {                                     // a NaN among the values (rare): the smallest disparity holding one
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (vx[k] != vx[k]) st.nani = min(st.nani, dtop - k);
    }
  }
}

// classify one chunk (warp-uniform) and scan it
template <bool EXT, bool DIV>
__device__ __forceinline__ void rr_any(RrState& st, const uint32_t (&v)[8], int c0w, int dtop, int D, int jfill, int llo, int lhi,
                                       float k2, float cnt) {
  if (c0w < jfill) rr_chunk8<EXT, DIV, 3>(st, v, c0w, dtop, D, jfill, k2, cnt);
  else if (c0w >= llo && c0w + 7 <= lhi) rr_chunk8<EXT, DIV, 0>(st, v, c0w, dtop, D, jfill, k2, cnt);
  else rr_chunk8<EXT, DIV, 4>(st, v, c0w, dtop, D, jfill, k2, cnt);
}

// fold state b into a (same pixel, disjoint disparities): equal extrema keep the smaller disparity
template <bool EXT>
__device__ __forceinline__ void rr_merge(RrState& a, float m2, float s2, float w2, float minv2, float maxv2, int mini2, int maxi2,
                                         int nani2, float k2) {
  const float M = fmaxf(a.m, m2);
  const float a1 = fast_exp2((a.m - M) * k2), a2 = fast_exp2((m2 - M) * k2);
  a.s = a.s * a1 + s2 * a2; a.ws = a.ws * a1 + w2 * a2; a.m = M;
  if (EXT) {
    if (minv2 < a.minv || (minv2 == a.minv && mini2 < a.mini)) { a.minv = minv2; a.mini = mini2; }
    if (maxv2 > a.maxv || (maxv2 == a.maxv && maxi2 < a.maxi)) { a.maxv = maxv2; a.maxi = maxi2; }
    a.nani = min(a.nani, nani2);
  }
}

template <bool EXT, bool DIV>
__global__ void __launch_bounds__(RR_THREADS, 1)
inner_regress_rows_kernel(RrOut out, RrGeom g, const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                          unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzled atoms: 1 KB alignment
  const uint32_t ab = (uint32_t)g.atom_bytes;
  unsigned char* sA = smem;                                    // nabuf x 2 atoms
  unsigned char* sB = sA + (size_t)g.nabuf * 2 * ab;           // nb atoms
  float* parts = reinterpret_cast<float*>(sB + (size_t)g.nb * ab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(parts) + RR_PART_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RR_MAXA + 2 * RR_MAXB + 4);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + 8 * RR_MAXA, b_full = a_empty + 8 * RR_MAXA,
                 b_empty = b_full + 8 * RR_MAXB, t_full = b_empty + 8 * RR_MAXB, t_empty = t_full + 8 * 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < RR_MAXA; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < RR_MAXB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, RR_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // the x-tiles of this CTA; consecutive tiles of one row share their right-window atoms
  const TileRange tr = tile_range(g);
  const uint32_t ntl = tr.ntl;
  const uint32_t na = (uint32_t)g.na, nb = (uint32_t)g.nb, nabuf = (uint32_t)g.nabuf, ng = (uint32_t)g.ng;

  if (warp == RR_EPI_WARPS) {
    // ================================================================ UMMA issuer
    // The whole warp runs the loop converged (everything below is warp-uniform: uniform registers) and ONE elected
    // lane issues; descriptor high words are constants, low words a base plus a multiple of 128 (2 KB per K step).
    // Branching on lane == 0 and rebuilding both 64-bit descriptors per MMA costs ~70 cycles per tcgen05.mma
    // (measured for the v4 volume kernel).
    if (ntl > 0) {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      const uint32_t idesc0 = (1u << 4) | ((uint32_t)g.fmt << 7) | ((uint32_t)g.fmt << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(RR_TM >> 4) << 24);     // + N >> 3 at bit 17
      // SWIZZLE_128B MN-major atoms as written by TMA: 8 channels x 64 pixels = 1 KB (SBO), the next 64 pixels one
      // atom further (LBO); a K = 16 step advances two channel groups (2 KB)
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO, descriptor version 1, SWIZZLE_128B
      const uint32_t lbo = ((ab >> 4) & 0x3FFFu) << 16;
      const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | lbo, b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFFu) | lbo;
      const uint32_t ab16 = ab >> 4;
      const int nks = g.C >> 4;
      long long c_op = 0, c_tm = 0;
      const long long c_beg = prof ? clock64() : 0;
      // ring positions kept incrementally (no divisions): left-tile slot / fill parity; slot / fill parity of the
      // window's first atom; how many atoms at the end of the window are new (not yet waited for); group counter
      uint32_t as = 0, apar = 0, bs0 = 0, bpar0 = 0, fresh = na, gc = 0;
      int xt = tr.first.xt;
      for (uint32_t tl = 0; tl < ntl; ++tl) {
        long long c0 = prof ? clock64() : 0;
        mbar_wait(a_full + 8 * as, apar);
        if (prof) c_op += clock64() - c0;
        const uint32_t a_lo = a_lo0 + as * 2 * ab16;
        uint32_t bs = bs0, bpar = bpar0, a = 0;
        for (uint32_t grp = 0; grp < ng; ++grp, ++gc) {
          const uint32_t gsz = grp == 0 ? (uint32_t)g.gsz0 : na - (uint32_t)g.gsz0;
          {                                                      // the group's new atoms have landed?
            c0 = prof ? clock64() : 0;
            uint32_t s = bs, p = bpar;
            for (uint32_t i = 0; i < gsz; ++i) {
              if (a + i + fresh >= na) mbar_wait(b_full + 8 * s, p);
              if (++s == nb) { s = 0; p ^= 1; }
            }
            if (prof) c_op += clock64() - c0;
          }
          const uint32_t half = gc & 1;
          c0 = prof ? clock64() : 0;
          mbar_wait(t_empty + 8 * half, ((gc >> 1) & 1) ^ 1);   // the epilogue has drained this half of TMEM
          if (prof) c_tm += clock64() - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t rem = gsz, col = half * 256;
          while (rem) {                                          // runs of contiguous ring slots (split only at the wrap)
            const uint32_t run = min(rem, nb - bs);
            if (leader) {
              const uint32_t td = tmem_base + col, b_lo = b_lo0 + bs * ab16, idesc = idesc0 | ((run * (RR_ATOM >> 3)) << 17);
              for (int ks = 0; ks < nks; ++ks)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                    "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                    ::"r"(td), "r"(a_lo + ks * 128), "r"(b_lo + ks * 128), "r"(DESC_HI), "r"(idesc), "r"(ks)
                    : "memory");
            }
            rem -= run; col += run * RR_ATOM; bs += run;
            if (bs == nb) { bs = 0; bpar ^= 1; }
          }
          if (leader) umma_commit(t_full + 8 * half);
          __syncwarp();
          a += gsz;
        }
        const bool cont = tl + 1 < ntl && xt + 1 < g.xtiles;    // the next tile continues this row
        const uint32_t nrel = cont ? 2u : na;                   // ... and keeps all but two atoms of this window
        if (leader) {
          umma_commit(a_empty + 8 * as);
          uint32_t rs = bs0;
          for (uint32_t i = 0; i < nrel; ++i) {
            umma_commit(b_empty + 8 * rs);
            if (++rs == nb) rs = 0;
          }
        }
        __syncwarp();
        bs0 += nrel; if (bs0 >= nb) { bs0 -= nb; bpar0 ^= 1; }
        fresh = nrel;
        if (++as == nabuf) { as = 0; apar ^= 1; }
        if (++xt == g.xtiles) xt = 0;
      }
      if (prof && leader) {
        atomicAdd(prof + 0, (unsigned long long)c_op);
        atomicAdd(prof + 1, (unsigned long long)c_tm);
        atomicAdd(prof + 2, (unsigned long long)(clock64() - c_beg));
      }
    }
  } else if (warp == RR_EPI_WARPS + 1) {
    // ================================================================ TMA producer of the right-atom ring (one lane)
    // (Two producer lanes in two warps, one per ring: a single lane issuing both in tile order was busy half of the
    // time at 64 channels -- ~350 cycles per box -- and a full ring stalled the requests for the other one.)
    if (lane == 0 && ntl > 0) {
      long long c_wait = 0;
      const long long c_beg = prof ? clock64() : 0;
      uint32_t bs = 0, bpar = 0;
      RowTile tc = tr.first;
      for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g)) {
        const bool first = tl == 0 || tc.xt == 0;
        const int x0 = tc.xt * RR_TM;
        for (uint32_t a = first ? 0u : na - 2u; a < na; ++a) {
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(b_empty + 8 * bs, bpar ^ 1);
          if (prof) c_wait += clock64() - c0;
          mbar_expect_tx(b_full + 8 * bs, ab);
          tma_load_4d(smem_u32(sB) + bs * ab, &tmR, b_full + 8 * bs, x0 - g.Dp + RR_ATOM * (int)a, tc.y, 0, tc.n);
          if (++bs == nb) { bs = 0; bpar ^= 1; }
        }
      }
      if (prof) {
        atomicAdd(prof + 3, (unsigned long long)c_wait);
        atomicAdd(prof + 4, (unsigned long long)(clock64() - c_beg));
      }
    }
  } else if (warp == RR_EPI_WARPS + 2) {
    // ================================================================ TMA producer of the left-tile ring (one lane)
    if (lane == 0 && ntl > 0) {
      uint32_t as = 0, apar = 0;
      RowTile tc = tr.first;
      for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g)) {
        const int x0 = tc.xt * RR_TM;
        mbar_wait(a_empty + 8 * as, apar ^ 1);
        const uint32_t dA = smem_u32(sA) + as * 2 * ab;
        mbar_expect_tx(a_full + 8 * as, 2 * ab);
        tma_load_4d(dA, &tmL, a_full + 8 * as, x0, tc.y, 0, tc.n);
        tma_load_4d(dA + ab, &tmL, a_full + 8 * as, x0 + RR_ATOM, tc.y, 0, tc.n);
        if (++as == nabuf) { as = 0; apar ^= 1; }
      }
    }
  } else {
    // ================================================================ epilogue
    const int q = warp & 3, hh = warp >> 2;
    const int r = 32 * q + lane;
    const bool pmean = g.mean && g.pow2;
    const float scale = pmean ? 1.f / (float)g.C : 1.f, cnt_f = (float)g.C;
    const float k2 = kLog2e * scale;
    const int D = g.D, Dp = g.Dp;
    const int dbw = r + Dp;                                     // disparity of this lane at window column j: dbw - j
    const int ulo = 32 * q + Dp - D + 1, uhi = 32 * q + 31 + Dp;   // columns any lane of the quadrant needs (inclusive)
    const int llo = ulo + 31, lhi = 32 * q + Dp;                  // columns every lane of the quadrant needs
    const int c8lo = ulo >> 3, c8hi = (uhi >> 3) + 1;             // ... as a range of 8-column chunks of the window
    const bool rec = prof && lane == 0;
    long long c_wait = 0, c_scan = 0, c_arr = 0, c_merge = 0;
    const long long c_beg = rec ? clock64() : 0;
    RrState st, st2;
    uint32_t gc = 0;
    int turn = 0;                                                // whose chunk is next among the quadrant's warps
    RowTile tc = tr.first;
    for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g)) {
      const int x0 = tc.xt * RR_TM, x = x0 + r;
      const int jfill = Dp - x0;                                 // window columns below this one have x' < 0
      st.reset(); st2.reset();
      int gat = 0;                                               // first atom of the group
      for (uint32_t grp = 0; grp < ng; ++grp, ++gc) {
        const int gsz = grp == 0 ? g.gsz0 : (int)na - g.gsz0;
        const uint32_t half = gc & 1;
        long long c0 = rec ? clock64() : 0;
        mbar_wait(t_full + 8 * half, (gc >> 1) & 1);
        if (rec) { const long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // this group's 8-column chunks inside the quadrant's range, every fourth one from this warp's turn on
        const int g8lo = max(c8lo, 8 * gat), g8hi = min(c8hi, 8 * (gat + gsz));
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + half * 256 - 64 * gat;   // + window column
        int c8 = g8lo + (hh - turn + RR_NSPLIT) % RR_NSPLIT;
        // two chunks per TMEM round trip, scanned into two independent states (st, st2) so that the two dependency chains
        // interleave -- a lone chunk is a ~250-cycle load + a ~250-cycle serial scan.  (Four chunks into four states, and
        // keeping the next pair's loads in flight during the scan, measured the same or slower.)
        for (; c8 + RR_NSPLIT < g8hi; c8 += 2 * RR_NSPLIT) {
          const int ca = 8 * c8, cb = ca + 8 * RR_NSPLIT;
          uint32_t va[8], vb[8];
          tmem_ld8(taddr + ca, va);
          tmem_ld8(taddr + cb, vb);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ca >= jfill && ca >= llo && cb + 7 <= lhi) {       // both in the interior of the band
            rr_chunk8<EXT, DIV, 0>(st, va, ca, dbw - ca, D, jfill, k2, cnt_f);
            rr_chunk8<EXT, DIV, 0>(st2, vb, cb, dbw - cb, D, jfill, k2, cnt_f);
          } else if (ca >= jfill) {
            rr_chunk8<EXT, DIV, 4>(st, va, ca, dbw - ca, D, jfill, k2, cnt_f);
            rr_chunk8<EXT, DIV, 4>(st2, vb, cb, dbw - cb, D, jfill, k2, cnt_f);
          } else {
            rr_chunk8<EXT, DIV, 3>(st, va, ca, dbw - ca, D, jfill, k2, cnt_f);
            rr_chunk8<EXT, DIV, 3>(st2, vb, cb, dbw - cb, D, jfill, k2, cnt_f);
          }
        }
        if (c8 < g8hi) {
          const int c0w = 8 * c8;
          uint32_t v[8];
          tmem_ld8(taddr + c0w, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          rr_any<EXT, DIV>(st, v, c0w, dbw - c0w, D, jfill, llo, lhi, k2, cnt_f);
        }
        if (g8hi > g8lo) turn = (turn + (g8hi - g8lo)) % RR_NSPLIT;
        if (rec) { const long long c1 = clock64(); c_scan += c1 - c0; c0 = c1; }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + 8 * half);
        if (rec) c_arr += clock64() - c0;
        gat += gsz;
      }
      const long long cm0 = rec ? clock64() : 0;
      // ---- merge: the warp's two states, then the four warps of the quadrant through shared memory (every warp parks
      // its state; the merging warp rotates with the tile so that no warp is the straggler every time)
      rr_merge<EXT>(st, st2.m, st2.s, st2.ws, st2.minv, st2.maxv, st2.mini, st2.maxi, st2.nani, k2);
      float* pbase = parts + (size_t)(tl & 1) * (RR_NSPLIT * 128 * 8);
      const int merger = (int)(tl % RR_NSPLIT);
      if (hh != merger) {
        float* p = pbase + ((size_t)hh * 128 + r) * 8;
        if (EXT) {
          *reinterpret_cast<float4*>(p) = make_float4(st.m, st.s, st.ws, st.minv);
          *reinterpret_cast<float4*>(p + 4) = make_float4(st.maxv, __int_as_float(st.mini), __int_as_float(st.maxi), __int_as_float(st.nani));
        } else {
          *reinterpret_cast<float4*>(p) = make_float4(st.m, st.s, st.ws, 0.f);
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "r"(32 * RR_NSPLIT) : "memory");
      if (hh == merger && x < g.W) {
#pragma unroll
        for (int k = 1; k < RR_NSPLIT; ++k) {
          const float* p = pbase + ((size_t)((merger + k) % RR_NSPLIT) * 128 + r) * 8;
          const float4 p0 = *reinterpret_cast<const float4*>(p);
          float4 p1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EXT) p1 = *reinterpret_cast<const float4*>(p + 4);
          rr_merge<EXT>(st, p0.x, p0.y, p0.z, p0.w, p1.x, __float_as_int(p1.y), __float_as_int(p1.z), __float_as_int(p1.w), k2);
        }
        const int64_t o = ((int64_t)tc.n * g.H + tc.y) * g.W + x;
        if (out.soft) out.soft[o] = __fdividef(st.ws, st.s);
        if (out.lse) out.lse[o] = st.m * scale + __logf(st.s);
        if (EXT) {
          if (st.nani != 0x7fffffff) { st.mini = st.nani; st.maxi = st.nani; }
          if (st.mini == 0x7fffffff) st.mini = 0;               // every value +inf: torch returns index 0
          if (st.maxi == 0x7fffffff) st.maxi = 0;
          if (out.amin) out.amin[o] = st.mini;
          if (out.amax) out.amax[o] = st.maxi;
        }
      }
      if (rec) c_merge += clock64() - cm0;
    }
    if (rec) {
      atomicAdd(prof + 7, (unsigned long long)c_scan);
      atomicAdd(prof + 8, (unsigned long long)c_arr);
      atomicAdd(prof + 9, (unsigned long long)c_merge);
      atomicAdd(prof + 5, (unsigned long long)c_wait);
      atomicAdd(prof + 6, (unsigned long long)(clock64() - c_beg));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// (W, H, C, N) view of a feature tensor; box = 64 pixels x 1 row x C channels, SWIZZLE_128B, zero fill outside
static bool rows_tmap(CUtensorMap* m, const rsm_feat& f, int fmt, const RrGeom& g, int64_t N) {
  const TmapEncodeFn enc = tmap_encoder();
  if (!enc || f.stride_w != 1 || !aligned_to(f.data, 16)) return false;
  const int64_t st[3] = {f.stride_h * 2, f.stride_c * 2, f.stride_n * 2};   // bytes
  const int64_t ext[3] = {g.H, g.C, N};
  cuuint64_t gstr[3];
  for (int i = 0; i < 3; ++i) {
    int64_t v = st[i];
    if (ext[i] == 1 && (v % 16 != 0 || v <= 0)) v = 16;                      // never stepped: any legal value
    if (v <= 0 || v % 16 != 0 || v >= (1LL << 40)) return false;
    gstr[i] = (cuuint64_t)v;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.C, (cuuint64_t)N};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {(cuuint32_t)RR_ATOM, 1, (cuuint32_t)g.C, 1};
  return enc(m, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(f.data),
             gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when this form does not apply (the caller falls back)
int launch_inner_regress_rows(const rsm_feat& left, const rsm_feat& right, int64_t N, int64_t C, int64_t H, int64_t W,
                              int64_t D, int mean, int in_dtype, const rsm_regress_out& out, cudaStream_t st,
                              unsigned long long* prof) {
  if (in_dtype != RSM_F16 && in_dtype != RSM_BF16) return RSM_ERR_UNSUPPORTED_CONFIG;
  if (C <= 0 || C % 16 != 0 || C > 128 || D <= 0 || D > 64 * (2 * RR_GATOMS - 2)) return RSM_ERR_UNSUPPORTED_CONFIG;
  RrGeom g;
  g.C = (int)C; g.H = (int)H; g.W = (int)W; g.D = (int)D;
  g.Dp = (int)ceil_div(D, RR_ATOM) * RR_ATOM;
  g.na = 2 + g.Dp / RR_ATOM;
  g.ng = g.na > RR_GATOMS ? 2 : 1;
  g.gsz0 = g.ng == 1 ? g.na : (g.na + 1) / 2;
  g.xtiles = (int)ceil_div(W, RR_TM);
  g.fmt = in_dtype == RSM_F16 ? 0 : 1;
  g.mean = mean;
  g.pow2 = (C & (C - 1)) == 0;
  g.atom_bytes = (int)C * 128;
  g.rows = N * H;
  g.tiles = g.rows * g.xtiles;
  if (g.tiles <= 0 || g.tiles > 2147483647LL) return RSM_ERR_UNSUPPORTED_CONFIG;
  const int64_t fixed = RR_PART_BYTES + RR_BAR_BYTES + 1024, ab = g.atom_bytes;
  g.nabuf = RR_MAXA;
  while (g.nabuf > 2 && fixed + g.nabuf * 2 * ab + (g.na + 2) * ab > RR_SMEM_MAX) --g.nabuf;
  int64_t nb = (RR_SMEM_MAX - fixed - g.nabuf * 2 * ab) / ab;
  if (nb > RR_MAXB) nb = RR_MAXB;
  if (nb < g.na + 1) return RSM_ERR_UNSUPPORTED_CONFIG;
  g.nb = (int)nb;
  alignas(64) CUtensorMap tmL, tmR;
  memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
  if (!rows_tmap(&tmL, left, g.fmt, g, N) || !rows_tmap(&tmR, right, g.fmt, g, N)) return RSM_ERR_UNSUPPORTED_CONFIG;
  const size_t smem = (size_t)(fixed + g.nabuf * 2 * ab + g.nb * ab);
  const RrOut ro{(float*)out.soft, out.argmin, out.argmax, out.lse};
  const bool ext = out.argmin || out.argmax, div = mean && !g.pow2;
  const unsigned grid = (unsigned)(g.tiles < kNumSMs ? g.tiles : kNumSMs);
  const char* where = "rsm_inner_regress_fwd(tcgen05 rows)";
  auto run = [&](auto kernel) -> int {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return finish_launch(where);
    kernel<<<grid, RR_THREADS, smem, st>>>(ro, g, tmL, tmR, prof);
    return finish_launch(where);
  };
  if (ext) return div ? run(inner_regress_rows_kernel<true, true>) : run(inner_regress_rows_kernel<true, false>);
  return div ? run(inner_regress_rows_kernel<false, true>) : run(inner_regress_rows_kernel<false, false>);
}

}  // namespace rsm
