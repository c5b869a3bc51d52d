// Adjoint of the inner-product / correlation volume on tcgen05 (16-bit tensors; 64 disparities per launch, see
// launch_inner_bwd_tc for D > 64):
//     gL[c, x]  = s * sum_d gV[d, x]      * R[c, x - d]        (x >= d)
//     gR[c, x'] = s * sum_d gV[d, x' + d] * L[c, x' + d]       (x' + d < W)
// -- what autograd derives from the slice-assign loop of TorchInnerProductCost.forward (cost_volume/inner_product.py:29-41)
// / make_correlation_volume (model/mobile_disp_net_c.py:188-205); SURVEY.md 8a "Backward contracts".  It is the forward's
// banded GEMM with the roles turned: per epipolar row and 128-pixel tile
//     gL_tile[128 x C] = A_L[128 x 192] . Rwin[192 x C],   A_L[r, j] = gV[r + 64 - j, x0 + r]   (0 <= r + 64 - j < D)
//     gR_tile[128 x C] = A_R[128 x 192] . Lwin[192 x C],   A_R[r, j] = gV[j - r, x0 + j]        (0 <= j - r < D)
// with Rwin = right pixels [x0 - 64, x0 + 128), Lwin = left pixels [x0, x0 + 192).  The band matrices A are what the
// SIMT kernels (rsm_corr_bwd.cuh) never build -- they walk the band with FFMAs, 0.11 of the HBM roofline in bf16 --
// here eight warps scatter the gradient tile into them (K-major SWIZZLE_128B atoms, 2-byte stores, conflict-free by
// the swizzle) and the contraction over the 192 window pixels runs on the tensor cores (M = 128, N = C block, K = 16;
// 3/4 of every MMA is zeros, which costs nothing that matters: the op is HBM-bound).
// Operands: the gradient tile gV[0..D) x [x0, x0 + 192) arrives as ONE TMA box (plain rows); feature windows are
// 64-pixel x C-block TMA boxes with SWIZZLE_128B -- as K-major UMMA operands a box row is one channel with 64 K
// elements -- streamed through rings exactly as in rsm_corr_rows.cu (consecutive tiles of a row share one atom).
// Warps: 0-7 band builders (a row per thread, half of the disparities each), 8-15 epilogue (TMEM lane quadrant = warp % 4,
// two parts sharing a tile's (gradient, 16-channel round) jobs: tcgen05.ld -> scale -> round -> staging tile -> TMA store),
// 16 UMMA issuer (converged warp, one elected lane), 17 TMA producer (one lane).  A_L and A_R are single buffers with their
// own ready / free barriers, so building one overlaps the MMAs on the other; two TMEM accumulator pairs decouple the MMAs
// from the epilogue.
#include <cuda.h>

#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int BT_TM = 128;                 // pixels per tile (UMMA M)
constexpr int BT_ATOM = 64;                // pixels per feature atom / K elements per band-matrix atom
constexpr int BT_DP = 64;                  // disparity reach of a window (D <= 64)
constexpr int BT_KATOMS = (BT_TM + BT_DP) / BT_ATOM;   // 3: K = 192 window pixels
constexpr int BT_GW = BT_TM + BT_DP;       // gradient tile width (pixels)
constexpr int BT_RING = 5;                 // slots per feature ring: window 3 + 2 ahead (4 where shared memory is short)
constexpr int BT_BUILD_WARPS = 8, BT_EPI_WARPS = 8;
constexpr int BT_THREADS = 32 * (BT_BUILD_WARPS + BT_EPI_WARPS + 2);
constexpr int BT_A_BYTES = BT_KATOMS * BT_TM * 128;    // one band matrix: 3 atoms of 128 rows x 128 bytes = 48 KB
constexpr int BT_BAR_BYTES = 512;         // mbarriers, the TMEM address slot and the builders' trash slot
constexpr int BT_SC = 16;                  // channels per output store
constexpr int BT_STILE = BT_SC * BT_TM * 2;             // one staging tile: 16 channel rows x 128 pixels
constexpr int BT_STAGE_BYTES = 4 * BT_STILE;           // two per epilogue part: the store of a round reads one while the next is written

struct BtGeom {
  int C, CB, cblocks;     // channels, channels per pass (<= 64, multiple of 16), passes
  int H, W, D, xtiles;
  int fmt;                // 0 = fp16, 1 = bf16
  int do_l, do_r;
  float scale;
  int d0;                 // first disparity of this launch's chunk (D = disparities of the chunk, <= 64)
  int gx;                 // x offset of the gradient box: 0, or d0 for the right gradient of a later chunk
  int rshift, lshift;     // feature windows: right pixels start at x0 - 64 - rshift, left pixels at x0 + lshift
  int accum;              // outputs are added to (TMA reduce-add), not stored: chunks after the first
  int atom_bytes;         // CB * 128
  int g_bytes;            // gradient tile buffer: D * 192 * 2 rounded up to 1 KB
  int tmem_cols;          // allocation: power of two >= 4 * CB
  int ring_r, ring_l;     // slots in use of the right- / left-feature ring (5, or 4)
  int64_t rows, tiles;    // N * H; cblocks * rows * xtiles
};

struct BtTile {
  int cb, n, y, xt;
  __device__ __forceinline__ void advance(const BtGeom& g, int N) {
    if (++xt < g.xtiles) return;
    xt = 0;
    if (++y < g.H) return;
    y = 0;
    if (++n < N) return;
    n = 0; ++cb;
  }
};

// byte offset of element (row r, k) inside a K-major SWIZZLE_128B band matrix: atoms of 64 k (128 rows x 128 B),
// 8-row groups 1 KB apart, the 16-byte chunk index XOR-ed with the row's position in its group
__device__ __forceinline__ uint32_t band_off(uint32_t rowbase, uint32_t rx16, int k) {
  return rowbase + (uint32_t)(k >> 6) * (BT_TM * 128) + ((((uint32_t)k << 1) & 126u) ^ rx16);
}

struct BtRing {   // slot / fill parity of a ring position
  uint32_t s, p, n;
  __device__ __forceinline__ void step(uint32_t k = 1) {
    s += k;
    if (s >= n) { s -= n; p ^= 1; }
  }
};

template <typename T, int ND>
__global__ void __launch_bounds__(BT_THREADS, 1)
inner_bwd_tc_kernel(BtGeom g, int N, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmL,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmGL,
                    const __grid_constant__ CUtensorMap tmGR, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t ab = (uint32_t)g.atom_bytes;
  unsigned char* sAL = smem;                                   // band matrix of the left gradient
  unsigned char* sAR = sAL + BT_A_BYTES;                       // ... of the right gradient
  unsigned char* sG = sAR + BT_A_BYTES;                        // 2 gradient tiles, rows of 192 pixels
  unsigned char* sRr = sG + 2 * (size_t)g.g_bytes;             // ring of right-feature atoms (operand of the left gradient)
  unsigned char* sLr = sRr + g.ring_r * (size_t)ab;            // ring of left-feature atoms
  unsigned char* sOut = sLr + g.ring_l * (size_t)ab;           // output staging tiles [channel][128 pixels]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + BT_STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t g_full = smem_u32(bars), g_empty = g_full + 16, al_ready = g_full + 32, al_free = g_full + 40,
                 ar_ready = g_full + 48, ar_free = g_full + 56, t_full = g_full + 64, t_empty = g_full + 80,
                 rr_full = g_full + 96, rr_empty = rr_full + 8 * BT_RING, lr_full = rr_empty + 8 * BT_RING,
                 lr_empty = lr_full + 8 * BT_RING;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(g_full + 8 * i, 1); mbar_init(g_empty + 8 * i, BT_BUILD_WARPS);
      mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, BT_EPI_WARPS);
    }
    mbar_init(al_ready, BT_BUILD_WARPS); mbar_init(al_free, 1);
    mbar_init(ar_ready, BT_BUILD_WARPS); mbar_init(ar_free, 1);
    for (int i = 0; i < BT_RING; ++i) {
      mbar_init(rr_full + 8 * i, 1); mbar_init(rr_empty + 8 * i, 1);
      mbar_init(lr_full + 8 * i, 1); mbar_init(lr_empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the band matrices start as zeros; every tile rewrites exactly the same band positions
  for (uint32_t o = 16u * threadIdx.x; o < 2u * BT_A_BYTES; o += 16u * BT_THREADS)
    *reinterpret_cast<uint4*>(sAL + o) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int64_t per = (g.tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_beg = min((int64_t)blockIdx.x * per, g.tiles), t_end = min(t_beg + per, g.tiles);
  const uint32_t ntl = (uint32_t)(t_end - t_beg);
  BtTile first;
  {
    uint32_t t = (uint32_t)t_beg;
    first.xt = (int)(t % (uint32_t)g.xtiles); t /= (uint32_t)g.xtiles;
    first.y = (int)(t % (uint32_t)g.H); t /= (uint32_t)g.H;
    first.n = (int)(t % (uint32_t)N);
    first.cb = (int)(t / (uint32_t)N);
  }
  const uint32_t CB = (uint32_t)g.CB;

  if (warp == BT_BUILD_WARPS + BT_EPI_WARPS) {
    // ================================================================ UMMA issuer (converged warp, one elected lane)
    if (ntl > 0) {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      // A and B K-major, fp32 accumulate, M = 128, N = CB
      const uint32_t idesc = (1u << 4) | ((uint32_t)g.fmt << 7) | ((uint32_t)g.fmt << 10) | ((CB >> 3) << 17) |
                             ((uint32_t)(BT_TM >> 4) << 24);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1 KB (8 rows), version 1, SWIZZLE_128B
      constexpr uint32_t LBO1 = 1u << 16;                                     // (unused for swizzled K-major operands)
      const uint32_t al_lo = ((smem_u32(sAL) >> 4) & 0x3FFFu) | LBO1, ar_lo = ((smem_u32(sAR) >> 4) & 0x3FFFu) | LBO1;
      const uint32_t rr_lo = ((smem_u32(sRr) >> 4) & 0x3FFFu) | LBO1, lr_lo = ((smem_u32(sLr) >> 4) & 0x3FFFu) | LBO1;
      const uint32_t ab16 = ab >> 4;
      BtRing r0{0, 0, (uint32_t)g.ring_r}, l0{0, 0, (uint32_t)g.ring_l};   // the windows' first atoms
      uint32_t fresh = BT_KATOMS;
      int xt = first.xt;
      long long c_t = 0, c_a = 0, c_ring = 0;
      const long long c_beg = prof ? clock64() : 0;
      auto side = [&](uint32_t a_lo, uint32_t ring_lo, uint32_t ring_full, BtRing w, uint32_t td) {
        for (uint32_t kb = 0; kb < BT_KATOMS; ++kb) {
          if (kb + fresh >= BT_KATOMS) {                                   // one of the window's new atoms
            const long long c0 = prof ? clock64() : 0;
            mbar_wait(ring_full + 8 * w.s, w.p);
            if (prof) c_ring += clock64() - c0;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint32_t a = a_lo + kb * ((BT_TM * 128) >> 4), b = ring_lo + w.s * ab16;
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              asm volatile(
                  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                  "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                  ::"r"(td), "r"(a + ks * 2), "r"(b + ks * 2), "r"(DESC_HI), "r"(idesc), "r"(kb | ks)
                  : "memory");
          }
          __syncwarp();
          w.step();
        }
      };
      for (uint32_t tl = 0; tl < ntl; ++tl) {
        const uint32_t buf = tl & 1, td = tmem_base + buf * 2 * CB;
        long long c0 = prof ? clock64() : 0;
        mbar_wait(t_empty + 8 * buf, ((tl >> 1) & 1) ^ 1);            // the epilogue has drained this accumulator pair
        if (prof) { const long long c1 = clock64(); c_t += c1 - c0; c0 = c1; }
        if (g.do_l) {
          mbar_wait(al_ready, tl & 1);
          if (prof) c_a += clock64() - c0;
          side(al_lo, rr_lo, rr_full, r0, td);
          if (leader) umma_commit(al_free);
        }
        if (g.do_r) {
          c0 = prof ? clock64() : 0;
          mbar_wait(ar_ready, tl & 1);
          if (prof) c_a += clock64() - c0;
          side(ar_lo, lr_lo, lr_full, l0, td + CB);
          if (leader) umma_commit(ar_free);
        }
        const bool cont = tl + 1 < ntl && xt + 1 < g.xtiles;
        const uint32_t nrel = cont ? 2u : (uint32_t)BT_KATOMS;
        if (leader) {
          umma_commit(t_full + 8 * buf);
          BtRing rr = r0, ll = l0;
          for (uint32_t i = 0; i < nrel; ++i, rr.step(), ll.step()) {
            if (g.do_l) umma_commit(rr_empty + 8 * rr.s);
            if (g.do_r) umma_commit(lr_empty + 8 * ll.s);
          }
        }
        __syncwarp();
        r0.step(nrel); l0.step(nrel);
        fresh = nrel;
        if (++xt == g.xtiles) xt = 0;
      }
      if (prof && leader) {
        atomicAdd(prof + 0, (unsigned long long)c_t);
        atomicAdd(prof + 1, (unsigned long long)c_a);
        atomicAdd(prof + 2, (unsigned long long)c_ring);
        atomicAdd(prof + 3, (unsigned long long)(clock64() - c_beg));
      }
    }
  } else if (warp == BT_BUILD_WARPS + BT_EPI_WARPS + 1) {
    // ================================================================ TMA producer (one lane)
    if (lane == 0 && ntl > 0) {
      BtRing rr{0, 0, (uint32_t)g.ring_r}, ll{0, 0, (uint32_t)g.ring_l};
      // the gradient tile runs one tile ahead of the feature atoms: its buffer frees up as soon as the builders are
      // done with the tile before last, while atoms wait for ring slots the MMAs of the previous tile still hold -- issued
      // in tile order, those waits kept the next gradient tile from being requested (builders 35 % idle waiting for it)
      auto load_g = [&](uint32_t tl, const BtTile& t) {
        const uint32_t gb = tl & 1;
        mbar_wait(g_empty + 8 * gb, ((tl >> 1) & 1) ^ 1);
        mbar_expect_tx(g_full + 8 * gb, (uint32_t)(g.D * BT_GW * 2));
        tma_load_4d(smem_u32(sG) + gb * (uint32_t)g.g_bytes, &tmG, g_full + 8 * gb, t.xt * BT_TM + g.gx, t.y, g.d0, t.n);
      };
      BtTile tc = first, tg = first;
      load_g(0, tg);
      for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g, N)) {
        const bool fst = tl == 0 || tc.xt == 0;
        const int x0 = tc.xt * BT_TM, c0 = tc.cb * (int)CB;
        if (tl + 1 < ntl) { tg.advance(g, N); load_g(tl + 1, tg); }
        for (uint32_t a = fst ? 0u : 1u; a < BT_KATOMS; ++a) {
          if (g.do_l) {
            mbar_wait(rr_empty + 8 * rr.s, rr.p ^ 1);
            mbar_expect_tx(rr_full + 8 * rr.s, ab);
            tma_load_4d(smem_u32(sRr) + rr.s * ab, &tmR, rr_full + 8 * rr.s, x0 - BT_DP - g.rshift + BT_ATOM * (int)a, tc.y, c0, tc.n);
          }
          if (g.do_r) {
            mbar_wait(lr_empty + 8 * ll.s, ll.p ^ 1);
            mbar_expect_tx(lr_full + 8 * ll.s, ab);
            tma_load_4d(smem_u32(sLr) + ll.s * ab, &tmL, lr_full + 8 * ll.s, x0 + g.lshift + BT_ATOM * (int)a, tc.y, c0, tc.n);
          }
          rr.step(); ll.step();
        }
      }
    }
  } else if (warp < BT_BUILD_WARPS) {
    // ================================================================ band builders
    // thread -> (row r, half of the disparities): ND disparities each (ND = the half rounded up to 8, a template
    // parameter so that everything below unrolls).  For a fixed d the 32 lanes of a warp read 64 contiguous bytes of the
    // gradient tile and scatter them along a diagonal of the band matrix (distinct 16-byte chunks by the swizzle).
    // Where an element goes depends on (r, d) only, never on the tile: the byte offsets are computed once and live in
    // registers, disparities past the end point at a trash slot -- per element the loop is one LDS with an immediate
    // offset and one STS (the first version recomputed the swizzled address and three predicates per element: ~15
    // instructions, 3500 cycles per tile, the bound of the kernel at 16 channels).
    // row of this thread: a warp takes every other 8-row group of its 64-row half (rows 8(2i + w%2) + l%8), which spreads
    // the scatter stores over the banks better than 32 consecutive rows (1.5 instead of 2 wavefronts per store)
    const int wq = (threadIdx.x >> 5) & 3, dh = threadIdx.x >> 7;
    const int r = 64 * (wq >> 1) + 16 * (lane >> 3) + 8 * (wq & 1) + (lane & 7);
    const int d_beg = dh ? (g.D + 1) / 2 : 0, d_end = dh ? g.D : (g.D + 1) / 2;
    const uint32_t rowbase = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u, rx16 = (uint32_t)(r & 7) << 4;
    const uint32_t trash = (uint32_t)(reinterpret_cast<unsigned char*>(bars) + 448 - smem);
    uint32_t offL[ND], offR[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      const int d = d_beg + i;
      offL[i] = d < d_end ? band_off(rowbase, rx16, r + BT_DP - d) : trash;
      offR[i] = d < d_end ? (uint32_t)BT_A_BYTES + band_off(rowbase, rx16, r + d) : trash;
    }
    const bool rec = prof && threadIdx.x == 0;
    long long c_g = 0, c_free = 0, c_build = 0;
    const long long c_beg = rec ? clock64() : 0;
    BtTile tc = first;
    for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g, N)) {
      const uint32_t gb = tl & 1;
      // row d_beg of the gradient tile at this thread's pixel (rows past D read whatever follows: they go to the trash slot)
      const unsigned short* G = reinterpret_cast<const unsigned short*>(sG + gb * (size_t)g.g_bytes) + d_beg * BT_GW + r;
      // x < d (the first tiles of a row; d counts from the chunk's first disparity): the forward never wrote that entry
      const int dmask = tc.xt * BT_TM + r - g.d0 - d_beg;
      long long c0 = rec ? clock64() : 0;
      mbar_wait(g_full + 8 * gb, (tl >> 1) & 1);
      if (rec) { const long long c1 = clock64(); c_g += c1 - c0; c0 = c1; }
      if (g.do_l) {
        mbar_wait(al_free, (tl & 1) ^ 1);                        // the MMAs of the previous tile have read A_L
        if (rec) { const long long c1 = clock64(); c_free += c1 - c0; c0 = c1; }
#pragma unroll
        for (int i0 = 0; i0 < ND; i0 += 8) {                     // eight loads in flight, then eight scattered stores
          unsigned short v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = G[(i0 + i) * BT_GW];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<unsigned short*>(smem + offL[i0 + i]) = i0 + i > dmask ? (unsigned short)0 : v[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(al_ready);
        if (rec) { const long long c1 = clock64(); c_build += c1 - c0; c0 = c1; }
      }
      if (g.do_r) {
        mbar_wait(ar_free, (tl & 1) ^ 1);
        if (rec) { const long long c1 = clock64(); c_free += c1 - c0; c0 = c1; }
        const unsigned short* Gd = G + d_beg;                    // gV[d, x' + d]: one more pixel per row
#pragma unroll
        for (int i0 = 0; i0 < ND; i0 += 8) {
          unsigned short v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = Gd[(i0 + i) * (BT_GW + 1)];
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<unsigned short*>(smem + offR[i0 + i]) = v[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(ar_ready);
        if (rec) c_build += clock64() - c0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(g_empty + 8 * gb);
    }
    if (rec) {
      atomicAdd(prof + 4, (unsigned long long)c_g);
      atomicAdd(prof + 5, (unsigned long long)c_free);
      atomicAdd(prof + 6, (unsigned long long)c_build);
      atomicAdd(prof + 7, (unsigned long long)(clock64() - c_beg));
    }
  } else {
    // ================================================================ epilogue
    // TMEM (lane = pixel, column = channel) -> scale -> round -> staging tile [channel][128 pixels] in shared memory
    // (a warp writes 64 contiguous bytes per channel) -> ONE TMA store per gradient and tile: whole 256-byte row
    // segments per channel, pixels past W clipped by the TMA unit.  (Storing straight from registers -- 2 bytes per lane,
    // 64-bit address arithmetic per channel -- kept the four epilogue warps 97 % busy and bounded the kernel at C = 64.)
    // Eight warps: TMEM lane quadrant q = warp % 4, part = which half of the (gradient, 16-channel round) jobs of a tile;
    // each part has its own staging tile, named barrier and storing lane.  (Four warps doing all of it were 95 % busy
    // and set the tile period at C = 64.)
    const int ew = warp - BT_BUILD_WARPS;
    const int q = ew & 3, part = ew >> 2;
    const int r = 32 * q + lane;
    const bool rec = prof && ew == 0 && lane == 0;
    const bool storer = q == 0 && lane == 0;
    long long c_w = 0;
    const long long c_beg = rec ? clock64() : 0;
    uint32_t nround = 0;                                        // staging rounds of this part so far
    BtTile tc = first;
    for (uint32_t tl = 0; tl < ntl; ++tl, tc.advance(g, N)) {
      const uint32_t buf = tl & 1;
      const long long c0 = rec ? clock64() : 0;
      mbar_wait(t_full + 8 * buf, (tl >> 1) & 1);
      if (rec) c_w += clock64() - c0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + buf * 2 * CB;
      int job = 0;
#pragma unroll 1
      for (int sd = 0; sd < 2; ++sd) {
        if (!(sd == 0 ? g.do_l : g.do_r)) continue;
#pragma unroll 1
        for (uint32_t cs = 0; cs < CB; cs += BT_SC, ++job) {       // 16 channels per staging round
          if ((job & 1) != part) continue;
          unsigned char* stile = sOut + (2 * part + (nround++ & 1)) * BT_STILE;
          T* stage = reinterpret_cast<T*>(stile) + r;
          if (storer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store before the previous one has read this tile
          asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
          uint32_t v[16];
          tmem_ld16(taddr + sd * CB + cs, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) stage[i * BT_TM] = from_f<T>(__uint_as_float(v[i]) * g.scale);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (TMA)
          asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
          if (storer) {
            if (!g.accum)
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                           ::"l"(sd == 0 ? &tmGL : &tmGR), "r"(smem_u32(stile)), "r"(tc.xt * BT_TM), "r"(tc.y),
                             "r"(tc.cb * (int)CB + (int)cs), "r"(tc.n)
                           : "memory");
            else   // a later disparity chunk: add to what the earlier chunks stored
              asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                           ::"l"(sd == 0 ? &tmGL : &tmGR), "r"(smem_u32(stile)), "r"(tc.xt * BT_TM), "r"(tc.y),
                             "r"(tc.cb * (int)CB + (int)cs), "r"(tc.n)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8 * buf);
    }
    if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");           // stores complete before the CTA exits
    if (rec) {
      atomicAdd(prof + 8, (unsigned long long)c_w);
      atomicAdd(prof + 9, (unsigned long long)(clock64() - c_beg));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
}

// 4-D tiled tensor map over a 16-bit tensor seen as (W, d1, d2, d3) with byte strides st[3]
static bool bt_tmap(CUtensorMap* m, const void* data, int fmt, const int64_t dims[4], const int64_t st[3], const uint32_t box[4],
                    bool swizzle) {
  const TmapEncodeFn enc = tmap_encoder();
  if (!enc || !aligned_to(data, 16)) return false;
  cuuint64_t gdim[4], gstr[3];
  for (int i = 0; i < 4; ++i) gdim[i] = (cuuint64_t)dims[i];
  for (int i = 0; i < 3; ++i) {
    int64_t v = st[i];
    if (dims[i + 1] == 1 && (v % 16 != 0 || v <= 0)) v = 16;                  // never stepped: any legal value
    if (v <= 0 || v % 16 != 0 || v >= (1LL << 40)) return false;
    gstr[i] = (cuuint64_t)v;
  }
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  return enc(m, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(data), gdim,
             gstr, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int ND>
static int bt_launch(const BtGeom& g, int N, const CUtensorMap& tmG, const CUtensorMap& tmL, const CUtensorMap& tmR,
                     const CUtensorMap& tmGL, const CUtensorMap& tmGR, unsigned grid, size_t smem, cudaStream_t st,
                     unsigned long long* prof) {
  const char* where = "rsm_inner_bwd(tcgen05)";
  auto kernel = inner_bwd_tc_kernel<T, ND>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return finish_launch(where);
  kernel<<<grid, BT_THREADS, smem, st>>>(g, N, tmG, tmL, tmR, tmGL, tmGR, prof);
  return finish_launch(where);
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when this form does not apply (the caller falls back to the SIMT kernels):
// gout (N,D,H,W), gl / gr (N,C,H,W) dense and of the features' 16-bit dtype, C a multiple of 16 and either <= 64 or a
// multiple of 64, D <= 512, rows TMA can address (W % 8 == 0).
// D > 64 runs as chunks of 64 disparities, one launch each (two from the second chunk on: the right gradient of chunk
// d0 reads the gradient box and the left window d0 pixels further right, the left gradient reads the right window d0
// pixels further left -- same kernel, shifted TMA coordinates); later chunks add to the stored result with TMA
// reduce-add.  Each launch re-reads the features (2 C H W) and its share of the gradient: at C = 128, D = 192 that is
// 2.2x the algorithmic traffic instead of the SIMT kernel's 0.04 of the HBM roofline.
int launch_inner_bwd_tc(const void* gout, const rsm_feat& left, const rsm_feat& right, void* gl, void* gr, int64_t N,
                        int64_t C, int64_t H, int64_t W, int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st,
                        unsigned long long* prof) {
  if ((in_dtype != RSM_F16 && in_dtype != RSM_BF16) || out_dtype != in_dtype) return RSM_ERR_UNSUPPORTED_CONFIG;
  if (C <= 0 || C % 16 != 0 || (C > 64 && C % 64 != 0) || D <= 0 || D > 8 * BT_DP || W % 8 != 0 || N <= 0 || H <= 0)
    return RSM_ERR_UNSUPPORTED_CONFIG;
  if (left.stride_w != 1 || right.stride_w != 1) return RSM_ERR_UNSUPPORTED_CONFIG;
  if ((gl && !aligned_to(gl, 16)) || (gr && !aligned_to(gr, 16))) return RSM_ERR_UNSUPPORTED_CONFIG;
  BtGeom g;
  g.C = (int)C; g.CB = C < 64 ? (int)C : 64; g.cblocks = (int)(C / g.CB);
  g.H = (int)H; g.W = (int)W;
  g.xtiles = (int)ceil_div(W, BT_TM);
  g.fmt = in_dtype == RSM_F16 ? 0 : 1;
  g.scale = mean ? 1.f / (float)C : 1.f;
  g.atom_bytes = g.CB * 128;
  g.tmem_cols = 32;
  while (g.tmem_cols < 4 * g.CB) g.tmem_cols *= 2;
  g.rows = N * H;
  g.tiles = (int64_t)g.cblocks * g.rows * g.xtiles;
  if (g.tiles > 2147483647LL) return RSM_ERR_UNSUPPORTED_CONFIG;
  alignas(64) CUtensorMap tmL, tmR, tmGL, tmGR;
  memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
  memset(&tmGL, 0, sizeof(tmGL)); memset(&tmGR, 0, sizeof(tmGR));
  auto feat_map = [&](CUtensorMap* m, const rsm_feat& f) {
    const int64_t dims[4] = {W, H, C, N}, strides[3] = {f.stride_h * 2, f.stride_c * 2, f.stride_n * 2};
    const uint32_t box[4] = {(uint32_t)BT_ATOM, 1, (uint32_t)g.CB, 1};
    return bt_tmap(m, f.data, g.fmt, dims, strides, box, true);
  };
  if (!feat_map(&tmL, left) || !feat_map(&tmR, right)) return RSM_ERR_UNSUPPORTED_CONFIG;
  const unsigned grid = (unsigned)(g.tiles < kNumSMs ? g.tiles : kNumSMs);
  // outputs: dense (N, C, H, W); box = 128 pixels x 16 channels, plain rows (the staging tile)
  auto out_map = [&](CUtensorMap* m, void* p) {
    if (!p) return true;
    const int64_t dims[4] = {W, H, C, N}, strides[3] = {W * 2, H * W * 2, C * H * W * 2};
    const uint32_t box[4] = {(uint32_t)BT_TM, 1, (uint32_t)(g.CB < BT_SC ? g.CB : BT_SC), 1};
    return bt_tmap(m, p, g.fmt, dims, strides, box, false);
  };
  if (!out_map(&tmGL, gl) || !out_map(&tmGR, gr)) return RSM_ERR_UNSUPPORTED_CONFIG;

  // one launch: disparities [d0, d0 + dc) into the gradients named by do_l / do_r
  auto chunk = [&](int d0, int dc, bool do_l, bool do_r) -> int {
    g.D = dc; g.d0 = d0;
    g.do_l = do_l; g.do_r = do_r;
    g.gx = do_r ? d0 : 0; g.lshift = d0; g.rshift = d0;
    g.accum = d0 > 0;
    g.g_bytes = (dc * BT_GW * 2 + 1023) / 1024 * 1024;
    const size_t fixed = 2 * (size_t)BT_A_BYTES + 2 * (size_t)g.g_bytes + BT_STAGE_BYTES + BT_BAR_BYTES + 1024;
    g.ring_r = g.ring_l = BT_RING;
    if (fixed + (size_t)(g.ring_r + g.ring_l) * g.atom_bytes > 227 * 1024) g.ring_l = BT_RING - 1;
    if (fixed + (size_t)(g.ring_r + g.ring_l) * g.atom_bytes > 227 * 1024) g.ring_r = BT_RING - 1;
    const size_t smem = fixed + (size_t)(g.ring_r + g.ring_l) * g.atom_bytes;
    if (smem > 227 * 1024) return (int)RSM_ERR_UNSUPPORTED_CONFIG;
    alignas(64) CUtensorMap tmG;
    memset(&tmG, 0, sizeof(tmG));
    const int64_t dims[4] = {W, H, D, N}, strides[3] = {W * 2, H * W * 2, D * H * W * 2};
    const uint32_t box[4] = {(uint32_t)BT_GW, 1, (uint32_t)dc, 1};
    if (!bt_tmap(&tmG, gout, g.fmt, dims, strides, box, false)) return (int)RSM_ERR_UNSUPPORTED_CONFIG;
    const int nd = ((dc + 1) / 2 + 7) / 8 * 8;                 // disparities per builder thread, rounded up to 8
    auto go = [&](auto tag) -> int {
      using T = decltype(tag);
      switch (nd) {
        case 8: return bt_launch<T, 8>(g, (int)N, tmG, tmL, tmR, tmGL, tmGR, grid, smem, st, prof);
        case 16: return bt_launch<T, 16>(g, (int)N, tmG, tmL, tmR, tmGL, tmGR, grid, smem, st, prof);
        case 24: return bt_launch<T, 24>(g, (int)N, tmG, tmL, tmR, tmGL, tmGR, grid, smem, st, prof);
        default: return bt_launch<T, 32>(g, (int)N, tmG, tmL, tmR, tmGL, tmGR, grid, smem, st, prof);
      }
    };
    if (in_dtype == RSM_F16) return go(__half{});
    return go(__nv_bfloat16{});
  };
  for (int d0 = 0; d0 < (int)D; d0 += BT_DP) {
    const int dc = (int)D - d0 < BT_DP ? (int)D - d0 : BT_DP;
    if (d0 == 0) {
      if (int rc = chunk(0, dc, gl != nullptr, gr != nullptr)) return rc;
    } else {
      if (gl) { if (int rc = chunk(d0, dc, true, false)) return rc; }
      if (gr) { if (int rc = chunk(d0, dc, false, true)) return rc; }
    }
  }
  return RSM_OK;
}

}  // namespace rsm
