// Inner-product / correlation VOLUME on the 5th-gen tensor cores (tcgen05), 16-bit features: the (N,D,H,W) volume of
// TorchInnerProductCost.forward (cost_volume/inner_product.py:11-42) / make_correlation_volume
// (model/mobile_disp_net_c.py:188-205).  (The fused build -> regression form lives in rsm_corr_rows.cu; fp32 features
// stay on the SIMT kernels of rsm_corr.cu: kind::tf32 keeps 10 mantissa bits, and the exact 3xTF32 split that was
// built and measured in round 1 never beat them -- it staged 2.5x the operand bytes and was removed.)
//
// Per epipolar row the correlation is the band  0 <= x - x' < D  of the W x W product
// P[x, x'] = sum_c L[c, x] * R[c, x']  (the reference's own einsum hint, cost_volume/inner_product.py:33-34).
// A tile = TM = 128 left pixels x0.. of one (n, y) and a chunk of DCH <= 128 disparities dc0..:
//     D_tmem[r, j] = sum_c L[c, x0 + r] * R[c, xr0 + j],   xr0 = x0 - dc0 - DCH,  j in [0, 128 + DCH)
// is ONE tcgen05.mma per 16 channels (M = 128, N = 128 + DCH, K = 16, both operands MN-major in shared
// memory, fp32 accumulators in TMEM, issued by one thread).  The value for disparity dc0 + dl of
// pixel x0 + r sits at column j = r + DCH - dl: a diagonal band.
//
// Persistent, warp-specialised CTA (448 threads, one per SM); each CTA owns a contiguous range of tiles:
//   warp 13    TMA producer (features whose strides TMA accepts -- the normal case): ONE lane arms smem_full[s] with
//              the stage's byte count and issues the cp.async.bulk.tensor box loads of a k-chunk (5-6 boxes of
//              64 pixels x <= 64 channels, SWIZZLE_128B): exactly the MN-major UMMA atoms, out-of-range pixels
//              zero-filled by the TMA unit.  Measured: the LSU path below tops out at ~2.9 TB/s of L2 -> shared
//              traffic on B200 whatever its pipeline depth.
//   warps 8-11 cp.async loaders (views TMA cannot address): stage the operand slab (<= 64 channels) of a k-chunk
//              into one of 2-6 shared-memory stages with 16-byte cp.async (features may be strided views; the
//              right window is zero-filled on both sides by the same instruction), writing the canonical
//              no-swizzle MN-major core-matrix layout directly:
//                  addr(x, c) = ((c/8) * (T/8) + x/8) * 128 + (c%8) * 16 + (x%8) * 2;
//              up to nstage-1 newer cp.async groups stay in flight behind the one being waited for.
//              On the TMA path they are a third epilogue group (disparity thirds).
//   warp 12    UMMA issuer (one lane): waits smem_full / tmem_empty, issues the tcgen05.mma chain of the
//              k-chunk, commits it to smem_empty (stage reusable) and tmem_full (accumulator ready);
//   warps 0-7  epilogue, two (three) warps per TMEM lane quadrant, each taking a part of the disparities:
//              tcgen05.ld the columns covering its lanes, park them in a padded shared-memory row per
//              lane, read them back skewed so that for every disparity 32 lanes hold 32 consecutive x, then
//              store the (N,D,H,W) volume with coalesced streaming stores.
// Two TMEM accumulator buffers decouple the UMMAs of tile t+1 from the epilogue of tile t.
// mbarriers: smem_full[s] (TMA bytes or 128 loader arrivals), smem_empty[s] (UMMA commit), tmem_full[b] (UMMA commit),
// tmem_empty[b] (epilogue arrivals).  All waits are bounded: a wait that expires traps (launch failure ->
// RSM_ERR_CUDA), never a hang, never a silently wrong result.
#include <cuda.h>   // CUtensorMap (types only; the encoder is looked up at run time, no libcuda link dependency)

#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int TC_TM = 128;        // UMMA M: left pixels per tile
constexpr int TC_KC = 64;         // channels per shared-memory stage (16-bit features)
constexpr int TC_NSTAGE = 6;      // upper bound on operand stages (g.nstage = 2..6, whatever fits in shared memory)
constexpr int TC_BAR_BYTES = 256; // mbarriers (2 per stage + 4 TMEM) and the TMEM address slot
constexpr int TC_EPI_WARPS = 8;    // warps 0-7: epilogue (two per TMEM lane quadrant)
constexpr int TC_THREADS = 32 * TC_EPI_WARPS + 128 + 64;   // + warps 8-11: loaders / splitters / third epilogue group,
                                                          //   warp 12: UMMA issuer, warp 13: TMA producer

struct TcGeom {
  int C, H, W, D;
  int dch;        // disparities per tile chunk (multiple of 16, <= 128)
  int ncol;       // UMMA N = TC_TM + dch
  int pitch;      // floats per lane row of the skew buffer
  int epi_bytes;  // epilogue scratch: skew rows
  int xtiles;     // ceil(W / TC_TM)
  int dchunks;    // ceil(D / dch)
  int mean, pow2;
  int fmt;        // 0 = fp16, 1 = bf16 (UMMA a/b format)
  int tmem_buf;   // TMEM columns per accumulator buffer (128 or 256)
  int stage_bytes;
  int nsplit;     // epilogue warps per TMEM lane quadrant (2, or 3 when warps 8-11 are free: 16-bit TMA volume path)
  int eb[4];      // disparity bounds of the epilogue parts inside a chunk: part p covers [eb[p], eb[p+1])
  int tma;        // operands arrive by TMA (SWIZZLE_128B atoms) instead of the cp.async loaders (no-swizzle atoms)
  int boxc;       // TMA: channels per box / per k-chunk (<= 64, multiple of 16)
  int nbb;        // TMA: 64-pixel boxes of the right window = ceil(ncol / 64)
  int nstage;     // operand stages in use: the loaders run up to nstage-1 k-chunks ahead of the UMMAs
  int64_t rows;   // N * H
  int64_t tiles;  // rows * xtiles * dchunks
};

template <typename Tin>
__device__ __forceinline__ uint4 load_chunk_slow(const Tin* __restrict__ src, int x, int W, int64_t sw) {
  constexpr int EPC = 16 / (int)sizeof(Tin);
  union { uint4 u; Tin e[EPC]; } tmp;
  tmp.u = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < EPC; ++i)
    if (x + i >= 0 && x + i < W) tmp.e[i] = __ldg(src + (int64_t)(x + i) * sw);
  return tmp.u;
}

// ---- stage nch channels of one operand with the 128 loader threads (lt = 0..127): xs = first x of the
// tile, nxg = x-groups of 16 bytes (8 or 4 elements).  thread -> (channel inside its K-group: 8 lanes write 128 contiguous bytes,
// x-group lane).  Chunks inside the image go global -> shared with 16-byte cp.async (LDGSTS: no register
// staging, every load of the stage in flight at once), chunks outside are zero-filled by the same
// instruction (src-size 0); ragged / unaligned chunks take the synchronous element-wise path.
template <typename Tin>
__device__ __forceinline__ void stage_operand(const FeatView& F, int64_t n, int y, int c0, int nch, int xs, int nxg, int W,
                                              unsigned char* dst, bool fast, int lt) {
  constexpr int EPC = 16 / (int)sizeof(Tin);
  const int cl = lt & 7, xl = lt >> 3, nxl = 16;
  const int ncg = nch >> 3;
  const Tin* __restrict__ base =
      reinterpret_cast<const Tin*>(F.data) + n * F.sn + (int64_t)y * F.sh + (int64_t)(c0 + cl) * F.sc;
  for (int xg = xl; xg < nxg; xg += nxl) {
    const int x = xs + EPC * xg;
    const bool inside = fast && x >= 0 && x + EPC <= W;
    const bool empty = x + EPC <= 0 || x >= W;
    unsigned char* d0 = dst + ((size_t)xg * 8 + cl) * 16;
    if (inside || empty) {
      const Tin* src = inside ? base + x : base;
      const int64_t cstep = inside ? 8 * F.sc : 0;
      const int nbytes = inside ? 16 : 0;
      for (int cg = 0; cg < ncg; ++cg)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(d0 + (size_t)cg * nxg * 128)),
                     "l"(src + cg * cstep), "r"(nbytes)
                     : "memory");
    } else {
      for (int cg = 0; cg < ncg; ++cg)
        *reinterpret_cast<uint4*>(d0 + (size_t)cg * nxg * 128) =
            load_chunk_slow<Tin>(base + (int64_t)(8 * cg) * F.sc, x, W, F.sw);
    }
  }
}

struct TileCoord {
  int64_t n;
  int y, x0, dc0;
  int xt;
  // next tile in (dchunk, n, y, xt) order, xt fastest: no divisions in the steady state
  __device__ __forceinline__ void advance(const TcGeom& g);
};
// 32-bit arithmetic only (tiles < 2^31 is checked on the host): 64-bit divisions are ~150-instruction
// dependent chains and this runs once per tile in every loader and epilogue warp
__device__ __forceinline__ TileCoord tile_coord(int64_t t64, const TcGeom& g) {
  TileCoord c;
  uint32_t t = (uint32_t)t64;
  const uint32_t xt = t % (uint32_t)g.xtiles; t /= (uint32_t)g.xtiles;
  const uint32_t rows = (uint32_t)g.rows;
  const uint32_t dchunk = t / rows, row = t - dchunk * rows;
  const uint32_t n = row / (uint32_t)g.H;
  c.n = n;
  c.y = (int)(row - n * (uint32_t)g.H);
  c.xt = (int)xt;
  c.x0 = (int)xt * TC_TM;
  c.dc0 = (int)dchunk * g.dch;
  return c;
}
__device__ __forceinline__ void TileCoord::advance(const TcGeom& g) {
  if (++xt < g.xtiles) { x0 += TC_TM; return; }
  xt = 0; x0 = 0;
  if (++y < g.H) return;
  y = 0;
  if (++n < g.rows / g.H) return;
  n = 0;
  dc0 += g.dch;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(TC_THREADS, 1)
inner_tc_kernel(FeatView L, FeatView R, Tout* __restrict__ out, TcGeom g, int fast,
                const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // swizzled atoms (TMA destinations, UMMA descriptors with base_offset 0) need 1024-byte alignment: 1 KB of slack
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int ES = (int)sizeof(Tin), EPC = 16 / ES;       // element bytes, elements per 16-byte chunk
  constexpr int KC = TC_KC;                                 // channels per stage
  unsigned char* stage0 = smem_raw;   // NSTAGE x { A: KC*128*ES | B: KC*ncol*ES }
  float* skew = reinterpret_cast<float*>(smem_raw + g.nstage * (size_t)g.stage_bytes);   // epilogue scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(skew) + g.epi_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NSTAGE + 4);
  const uint32_t smem_empty = smem_u32(bars), smem_full = smem_u32(bars + TC_NSTAGE),
                 tmem_full = smem_u32(bars + 2 * TC_NSTAGE), tmem_empty = smem_u32(bars + 2 * TC_NSTAGE + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * g.tmem_buf))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_NSTAGE; ++i) {
      mbar_init(smem_empty + 8 * i, 1);        // one UMMA commit
      mbar_init(smem_full + 8 * i, g.tma ? 1 : 128);   // TMA: the producer's expect_tx; else every loader thread
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 32 * 4 * g.nsplit);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int nk = g.tma ? (g.C + g.boxc - 1) / g.boxc : (g.C + KC - 1) / KC;
  // contiguous tile range of this CTA (neighbouring x tiles share most of their right window in L2)
  const int64_t per = (g.tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_beg = min((int64_t)blockIdx.x * per, g.tiles), t_end = min(t_beg + per, g.tiles);

  if (warp == TC_EPI_WARPS + 4) {
    // ================================================================ UMMA issuer (one elected lane)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)g.fmt << 7) | ((uint32_t)g.fmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(g.ncol >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
      const uint32_t sbo = 128, lboA = (TC_TM / EPC) * 128, lboB = (uint32_t)(g.ncol / EPC) * 128;
      const uint32_t nst = (uint32_t)g.nstage;
      uint32_t it = 0, use = 0;
      for (int64_t t = t_beg; t < t_end; ++t, ++use) {
        const uint32_t buf = use & 1;
        mbar_wait(tmem_empty + 8 * buf, ((use >> 1) & 1) ^ 1);            // epilogue drained this accumulator
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const uint32_t s = it % nst;
          const uint32_t sA = smem_u32(stage0 + (size_t)s * g.stage_bytes), sB = sA + KC * TC_TM * ES;
          const int nch = min(KC, g.C - kc * KC);
          mbar_wait(smem_full + 8 * s, (it / nst) & 1);                   // operands of this k-chunk have landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t td = tmem_base + buf * g.tmem_buf;
          if (g.tma) {
            // SWIZZLE_128B atoms as written by TMA: 8 channels x 64 pixels (1 KB); next 8 channels +1 KB (SBO),
            // next 64 pixels one box further (LBO); a K = 16 UMMA advances two channel groups
            const uint32_t boxb = (uint32_t)g.boxc * 128u, sBt = sA + 2 * boxb;
            for (int ks = 0; ks < g.boxc / 16; ++ks) {
              const uint64_t adesc = umma_desc(sA + ks * 2048, boxb, 1024, 2);
              const uint64_t bdesc = umma_desc(sBt + ks * 2048, boxb, 1024, 2);
              umma_f16(td, adesc, bdesc, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
            }
          } else {
            for (int ks = 0; ks < nch / 16; ++ks) {                       // K = 16 per UMMA: two K-groups
              const uint64_t adesc = umma_desc(sA + ks * 2 * lboA, lboA, sbo);
              const uint64_t bdesc = umma_desc(sB + ks * 2 * lboB, lboB, sbo);
              umma_f16(td, adesc, bdesc, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_empty + 8 * s);                                // stage reusable once these complete
          if (kc == nk - 1) umma_commit(tmem_full + 8 * buf);             // accumulator ready for the epilogue
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 5) {
    // ============================================================================== TMA producer
    // One lane keeps every free stage loading: the stage-free barrier (UMMA commit) is the only throttle.
    if (lane == 0 && g.tma) {
      const uint32_t nst = (uint32_t)g.nstage;
      const uint32_t boxb = (uint32_t)g.boxc * 128u, bytes = (2u + (uint32_t)g.nbb) * boxb;
      uint32_t it = 0;
      TileCoord tc = tile_coord(t_beg, g);
      for (int64_t t = t_beg; t < t_end; ++t, tc.advance(g)) {
        const int xr0 = tc.x0 - tc.dc0 - g.dch;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const uint32_t s = it % nst;
          const uint32_t sA = smem_u32(stage0 + (size_t)s * g.stage_bytes), bar = smem_full + 8 * s;
          mbar_wait(smem_empty + 8 * s, ((it / nst) & 1) ^ 1);       // UMMAs that read this stage have completed
          mbar_expect_tx(bar, bytes);
          const int c0 = kc * g.boxc;
          tma_load_4d(sA, &tmL, bar, tc.x0, tc.y, c0, (int)tc.n);
          tma_load_4d(sA + boxb, &tmL, bar, tc.x0 + 64, tc.y, c0, (int)tc.n);
          for (int m = 0; m < g.nbb; ++m)
            tma_load_4d(sA + (2 + m) * boxb, &tmR, bar, xr0 + 64 * m, tc.y, c0, (int)tc.n);
        }
      }
    }
  } else if (warp >= TC_EPI_WARPS && g.tma && g.nsplit == 2) {
    // warps 8-11 have nothing to do (TMA path with the two-way epilogue)
  } else if (warp >= TC_EPI_WARPS && !g.tma) {
    // ================================================================================== loaders
    // One cp.async group per k-chunk job; up to nstage-1 newer groups stay in flight while the loaders
    // wait for the oldest one to land, fence it for the async proxy and signal smem_full.  The loaders
    // never wait on the UMMA issuer except for a free stage (smem_empty), nstage jobs later.
    const int lt = threadIdx.x - 32 * TC_EPI_WARPS;
    const uint32_t nst = (uint32_t)g.nstage;
    auto landed = [&](uint32_t job, int newer_groups_in_flight) {
      switch (newer_groups_in_flight) {   // wait_group takes an immediate
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      mbar_arrive(smem_full + 8 * (job % nst));
    };
    uint32_t it = 0, done = 0;   // k-chunk jobs issued / signalled so far
    TileCoord tc = tile_coord(t_beg, g);
    for (int64_t t = t_beg; t < t_end; ++t, tc.advance(g)) {
      const int xr0 = tc.x0 - tc.dc0 - g.dch;
      for (int kc = 0; kc < nk; ++kc, ++it) {
        const uint32_t s = it % nst;
        unsigned char* sA = stage0 + (size_t)s * g.stage_bytes;
        unsigned char* sB = sA + KC * TC_TM * ES;
        const int c0 = kc * KC, nch = min(KC, g.C - c0);
        mbar_wait(smem_empty + 8 * s, ((it / nst) & 1) ^ 1);       // UMMAs that read this stage have completed
        stage_operand<Tin>(L, tc.n, tc.y, c0, nch, tc.x0, TC_TM / EPC, g.W, sA, fast, lt);
        stage_operand<Tin>(R, tc.n, tc.y, c0, nch, xr0, g.ncol / EPC, g.W, sB, fast, lt);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it + 1 - done == nst) {             // keep at most nstage-1 newer groups behind the oldest
          landed(done, (int)nst - 1);
          ++done;
        }
      }
    }
    for (; done < it; ++done) landed(done, (int)(it - done - 1));   // drain
  } else {
    // ================================================================================= epilogue
    // warp -> (TMEM lane quadrant q, disparity part hh): lanes 32q.., disparities [eb[hh], eb[hh+1]) of the chunk
    // (halves; thirds when warps 8-11 join in: TMA path).  Disparity dl of lane t sits at column
    // 32q + t + dch - dl; this warp's window starts at 32q + cs with cs = dch - eb[hh+1] and is ncw columns wide
    // (>= part width + 32, multiple of 16).
    const int q = warp & 3, hh = warp >> 2;
    const int plo = hh == 0 ? g.eb[0] : hh == 1 ? g.eb[1] : g.eb[2];     // (static indices: no local copy of g)
    const int phi = hh == 0 ? g.eb[1] : hh == 1 ? g.eb[2] : g.eb[3];
    const int cs = g.dch - phi, ncw = (phi - plo + 32 + 15) / 16 * 16;
    const float inv = 1.f / (float)g.C, cnt = (float)g.C;
    uint32_t use = 0;
    TileCoord tc = tile_coord(t_beg, g);
    for (int64_t t = t_beg; t < t_end; ++t, ++use, tc.advance(g)) {
      const uint32_t buf = use & 1;
      mbar_wait(tmem_full + 8 * buf, (use >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + buf * g.tmem_buf + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * q + cs);
      const int x = tc.x0 + 32 * q + lane;
      const int dmax = min(g.dch, g.D - tc.dc0);
      const int dlo = plo, dhi = min(phi, dmax);

      // ---- TMEM -> one padded shared-memory row per lane (up to four 16-column loads in flight per wait)
      float* row = skew + (size_t)(32 * warp + lane) * g.pitch;
      for (int cb = 0; cb < ncw; cb += 64) {
        uint32_t r[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cb + 16 * u < ncw) tmem_ld16(taddr + cb + 16 * u, r[u]);   // warp-uniform condition
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cb + 16 * u < ncw) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<uint4*>(row + cb + 16 * u + i) = make_uint4(r[u][i], r[u][i + 1], r[u][i + 2], r[u][i + 3]);
          }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_empty + 8 * buf);                   // this thread is done with the TMEM buffer
      __syncwarp();
      // ---- skewed read-back: value(dl) = rp0[-dl]; zeros where x < d, i.e. for dl >= dz
      const float* rp0 = row + lane + phi;
      const float mul = g.mean ? (g.pow2 ? inv : 1.f) : 1.f;
      const bool divide = g.mean && !g.pow2;
      const int dz = max(dlo, min(dhi, x - tc.dc0 + 1));    // [dlo, dz): values, [dz, dhi): fill
      if (x < g.W) {
        const int64_t dstride = (int64_t)g.H * g.W;
        Tout* __restrict__ o = out + (((int64_t)tc.n * g.D + tc.dc0 + dlo) * g.H + tc.y) * g.W + x;
        // one uniform loop over the warp's disparities: the x < d fill is a select, not a second loop with a
        // per-lane trip count (on the first tile of a row the divergent version made three warps of the CTA 2-3x
        // slower than the rest, and the slowest epilogue warp sets the tile period -- measured with clock64)
        int dl = dlo;
        for (; dl + 4 <= dhi; dl += 4) {                   // 4 independent LDS -> STG chains
          float v0 = rp0[-dl], v1 = rp0[-dl - 1], v2 = rp0[-dl - 2], v3 = rp0[-dl - 3];
          if (divide) { v0 = v0 * mul / cnt; v1 = v1 * mul / cnt; v2 = v2 * mul / cnt; v3 = v3 * mul / cnt; }
          else { v0 *= mul; v1 *= mul; v2 *= mul; v3 *= mul; }
          __stcs(o, from_f<Tout>(dl < dz ? v0 : 0.f)); o += dstride;
          __stcs(o, from_f<Tout>(dl + 1 < dz ? v1 : 0.f)); o += dstride;
          __stcs(o, from_f<Tout>(dl + 2 < dz ? v2 : 0.f)); o += dstride;
          __stcs(o, from_f<Tout>(dl + 3 < dz ? v3 : 0.f)); o += dstride;
        }
        for (; dl < dhi; ++dl, o += dstride) {
          const float v = divide ? rp0[-dl] * mul / cnt : rp0[-dl] * mul;
          __stcs(o, from_f<Tout>(dl < dz ? v : 0.f));
        }
      }
      __syncwarp();                                         // rows are reused by the next tile
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * g.tmem_buf))
                 : "memory");
}

static int tc_geom(int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int mean, int fmt, TcGeom& g) {
  g.C = (int)C; g.H = (int)H; g.W = (int)W; g.D = (int)D;
  const int d16 = (int)((D + 15) / 16 * 16);
  g.dch = d16 < 128 ? d16 : 128;
  g.ncol = TC_TM + g.dch;
  g.xtiles = (int)ceil_div(W, TC_TM);
  g.dchunks = (int)ceil_div(D, g.dch);
  g.mean = mean;
  g.pow2 = (C & (C - 1)) == 0;
  g.fmt = fmt;
  g.tmem_buf = g.ncol <= 128 ? 128 : 256;
  g.stage_bytes = TC_KC * (TC_TM + g.ncol) * 2;
  g.rows = N * H;
  g.tiles = g.rows * g.xtiles * g.dchunks;
  if (g.tiles <= 0 || g.tiles > 2147483647LL) return RSM_ERR_INVALID_SHAPE;
  return RSM_OK;
}

// rows start on 16-byte boundaries: strides are multiples of epc = 16 / sizeof(element) elements
static bool feat_vec16(const rsm_feat& f, int epc) {
  return f.stride_w == 1 && f.stride_n % epc == 0 && f.stride_c % epc == 0 && f.stride_h % epc == 0 && aligned_to(f.data, 16);
}

// ---- tensor maps: (W, H, C, N) view of a feature tensor, box = 64 pixels x 1 row x boxc channels, SWIZZLE_128B
static bool make_tmap(CUtensorMap* m, const rsm_feat& f, int fmt, const TcGeom& g, int64_t N) {
  const TmapEncodeFn enc = tmap_encoder();
  if (!enc || f.stride_w != 1 || !aligned_to(f.data, 16)) return false;
  const int64_t st[3] = {f.stride_h * 2, f.stride_c * 2, f.stride_n * 2};   // bytes
  const int64_t ext[3] = {g.H, g.C, N};
  cuuint64_t gstr[3];
  for (int i = 0; i < 3; ++i) {
    int64_t v = st[i];
    if (ext[i] == 1 && (v % 16 != 0 || v <= 0)) v = 16;                        // never stepped: any legal value
    if (v <= 0 || v % 16 != 0 || v >= (1LL << 40)) return false;
    gstr[i] = (cuuint64_t)v;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.C, (cuuint64_t)N};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)g.boxc, 1};
  return enc(m, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(f.data),
             gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename Tin, typename Tout>
static int launch_tc(const rsm_feat& left, const rsm_feat& right, void* out, const TcGeom& g_in, cudaStream_t st,
                     const char* where) {
  TcGeom g = g_in;
  // operands by TMA when the views qualify, else the cp.async loaders
  alignas(64) CUtensorMap tmL, tmR;
  memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
  g.tma = 0;
  g.boxc = g.C < TC_KC ? g.C : TC_KC;
  g.nbb = (g.ncol + 63) / 64;
  if (make_tmap(&tmL, left, g.fmt, g, g.rows / g.H) && make_tmap(&tmR, right, g.fmt, g, g.rows / g.H)) {
    g.tma = 1;
    g.stage_bytes = (2 + g.nbb) * g.boxc * 128;
  }
  // epilogue parts: halves of the chunk, or thirds (bounds on multiples of 8) when warps 8-11 are free and the
  // wider scratch still leaves room for two operand stages
  auto epilogue_parts = [&](int nsplit) {
    g.nsplit = nsplit;
    auto r8 = [&](int v) { v = (v + 7) / 8 * 8; return v < g.dch ? v : g.dch; };
    g.eb[0] = 0;
    if (nsplit == 2) { g.eb[1] = g.dch / 2; g.eb[2] = g.dch; g.eb[3] = g.dch; }
    else { g.eb[1] = r8(g.dch / 3); g.eb[2] = r8(2 * g.dch / 3); g.eb[3] = g.dch; }
    int ncw = 0;                           // TMEM columns the widest part pulls: >= width + 32, multiple of 16
    for (int k = 0; k < nsplit; ++k) {
      const int c = (g.eb[k + 1] - g.eb[k] + 32 + 15) / 16 * 16;
      ncw = c > ncw ? c : ncw;
    }
    int p = ncw;                           // pitch: >= ncw, multiple of 4 with an odd quotient (conflict-free
    if ((p / 4) % 2 == 0) p += 4;          // 128-bit row writes and conflict-free skewed 32-bit reads)
    g.pitch = p;
    g.epi_bytes = 32 * 4 * nsplit * g.pitch * 4;
  };
  epilogue_parts(g.tma ? 3 : 2);
  if (g.nsplit == 3 && 2 * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024 > 220 * 1024) epilogue_parts(2);
  g.nstage = TC_NSTAGE;
  while (g.nstage > 2 && g.nstage * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024 > 220 * 1024) --g.nstage;
  const size_t smem = g.nstage * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024;
  auto k = inner_tc_kernel<Tin, Tout>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return finish_launch(where);
  const int epc = 16 / (int)sizeof(Tin);
  const int fast = feat_vec16(left, epc) && feat_vec16(right, epc);   // 16-byte chunks start at multiples of epc elements
  const unsigned grid = (unsigned)(g.tiles < kNumSMs ? g.tiles : kNumSMs);   // persistent: one CTA per SM
  k<<<grid, TC_THREADS, smem, st>>>(view_of(left), view_of(right), (Tout*)out, g, fast, tmL, tmR);
  return finish_launch(where);
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when the tensor-core path does not apply (caller falls back to SIMT):
// 16-bit features, whole UMMAs of K = 16
int launch_inner_tc(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                    int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st) {
  if (in_dtype == RSM_F32 || C <= 0 || D <= 0 || C % 16 != 0) return RSM_ERR_UNSUPPORTED_CONFIG;
  TcGeom g;
  if (int rc = tc_geom(N, C, H, W, D, mean, in_dtype == RSM_F16 ? 0 : 1, g)) return rc;
  const char* where = "rsm_inner_fwd(tcgen05)";
  if (in_dtype == RSM_F16) {
    if (out_dtype == RSM_F32) return launch_tc<__half, float>(left, right, out, g, st, where);
    return launch_tc<__half, __half>(left, right, out, g, st, where);
  }
  if (out_dtype == RSM_F32) return launch_tc<__nv_bfloat16, float>(left, right, out, g, st, where);
  return launch_tc<__nv_bfloat16, __nv_bfloat16>(left, right, out, g, st, where);
}

}  // namespace rsm
