// Inner-product / correlation on the 5th-gen tensor cores (tcgen05): the (N,D,H,W) volume (EPI_VOLUME)
// or, fused, the soft-argmax / argmin / argmax of it without ever writing the volume (EPI_REGRESS).
// fp16 / bf16 features use kind::f16; fp32 features use kind::tf32 three times per k-step on an exact
// hi/lo split of the operands (x = hi + lo, hi = top 11 significand bits): hi*hi + hi*lo + lo*hi with
// fp32 accumulation drops only the lo*lo term (~2^-22 relative), i.e. fp32-grade results ("3xTF32";
// opt-in, RSM_TC_FP32=1).  Non-finite fp32 inputs yield non-finite outputs at the same positions as the
// reference, but +-inf may come out as NaN (inf * lo with lo = 0).
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
//   warp 13    TMA producer (16-bit features whose strides TMA accepts -- the normal case -- and the opt-in fp32
//              path): ONE lane arms smem_full[s] / raw_full[s] with the stage's byte count and issues the
//              cp.async.bulk.tensor box loads of a k-chunk (16-bit: 5-6 boxes of 64 pixels x <= 64 channels,
//              SWIZZLE_128B; fp32: 10 boxes of 32 pixels x 16 channels, SWIZZLE_128B_ATOM_32B): exactly the
//              MN-major UMMA atoms, out-of-range pixels zero-filled by the TMA unit.  Measured: the LSU path below
//              tops out at ~2.9 TB/s of L2 -> shared traffic on B200 whatever its pipeline depth.
//   warps 8-11 cp.async loaders (views TMA cannot address): stage the operand slab (<= 64 channels) of a k-chunk
//              into one of 2-6 shared-memory stages with 16-byte cp.async (features may be strided views; the
//              right window is zero-filled on both sides by the same instruction), writing the canonical
//              no-swizzle MN-major core-matrix layout directly:
//                  addr(x, c) = ((c/8) * (T/8) + x/8) * 128 + (c%8) * 16 + (x%8) * 2;
//              up to nstage-1 newer cp.async groups stay in flight behind the one being waited for.
//              fp32 TMA path: the same warps split every landed raw chunk into hi (in place) and lo (one of two
//              rotating slots).  16-bit TMA volume path: they are a third epilogue group (disparity thirds).
//   warp 12    UMMA issuer (one lane): waits smem_full / tmem_empty, issues the tcgen05.mma chain of the
//              k-chunk, commits it to smem_empty (stage reusable) and tmem_full (accumulator ready);
//   warps 0-7  epilogue, two (three) warps per TMEM lane quadrant, each taking a part of the disparities:
//              tcgen05.ld the columns covering its lanes, park them in a padded shared-memory row per
//              lane, read them back skewed so that for every disparity 32 lanes hold 32 consecutive x, then
//              either store the (N,D,H,W) volume (EPI_VOLUME) or run the chunked online softmax +
//              arg-extrema over them (EPI_REGRESS; the two halves of a quadrant merge through smem).
// Two TMEM accumulator buffers decouple the UMMAs of tile t+1 from the epilogue of tile t.
// mbarriers: smem_full[s] (TMA bytes or 128 loader / splitter arrivals), smem_empty[s] (UMMA commit), raw_full[s]
// (fp32 TMA bytes), lo_empty[2] (UMMA commit), tmem_full[b] (UMMA commit), tmem_empty[b] (epilogue arrivals).
// All waits are bounded: a wait that expires traps (launch failure -> RSM_ERR_CUDA), never a hang, never a
// silently wrong result.
#include <cuda.h>   // CUtensorMap (types only; the encoder is looked up at run time, no libcuda link dependency)

#include "rsm_common.cuh"
#include "rsm_tc.cuh"

namespace rsm {

constexpr int TC_TM = 128;        // UMMA M: left pixels per tile
constexpr int TC_KC = 64;         // channels per shared-memory stage (16-bit features)
constexpr int TC_KC32 = 16;       // channels per stage for fp32 features (hi + lo copies: same stage bytes)
constexpr int TC_NSTAGE = 6;      // upper bound on operand stages (g.nstage = 2..6, whatever fits in shared memory)
constexpr int TC_BAR_BYTES = 256; // mbarriers (3 per stage + 4 TMEM + 2 lo-slot) and the TMEM address slot
constexpr int TC_EPI_WARPS = 8;    // warps 0-7: epilogue (two per TMEM lane quadrant)
constexpr int TC_THREADS = 32 * TC_EPI_WARPS + 128 + 64;   // + warps 8-11: loaders / splitters / third epilogue group,
                                                          //   warp 12: UMMA issuer, warp 13: TMA producer
enum { EPI_VOLUME = 0, EPI_REGRESS = 1 };

struct TcGeom {
  int C, H, W, D;
  int dch;        // disparities per tile chunk (multiple of 16, fp32: 32; <= 128)
  int ncol;       // UMMA N = TC_TM + dch
  int pitch;      // floats per lane row of the skew buffer
  int epi_bytes;  // epilogue scratch: skew rows (volume) or partial softmax states (regress)
  int xtiles;     // ceil(W / TC_TM)
  int dchunks;    // ceil(D / dch)
  int mean, pow2;
  int fmt;        // 0 = fp16, 1 = bf16, 2 = tf32 (UMMA a/b format)
  int tmem_buf;   // TMEM columns per accumulator buffer (128 or 256)
  int stage_bytes;
  int d_fastest;  // tile order: disparity chunk fastest (fused regress keeps per-pixel state across chunks)
  int nsplit;     // epilogue warps per TMEM lane quadrant (2, or 3 when warps 8-11 are free: 16-bit TMA volume path)
  int eb[4];      // disparity bounds of the epilogue parts inside a chunk: part p covers [eb[p], eb[p+1])
  int tma32;      // fp32 operands arrive by TMA into a deep raw/hi ring; splitter warps write hi in place and lo
                  // into one of two lo slots (stage_bytes = one hi stage; the lo slots follow the ring)
  int tma;        // operands arrive by TMA (SWIZZLE_128B atoms) instead of the cp.async loaders (no-swizzle atoms)
  int boxc;       // TMA: channels per box / per k-chunk (<= 64, multiple of 16)
  int nbb;        // TMA: 64-pixel boxes of the right window = ceil(ncol / 64)
  int nstage;     // operand stages in use: the loaders run up to nstage-1 k-chunks ahead of the UMMAs
  int64_t rows;   // N * H
  int64_t tiles;  // rows * xtiles * dchunks
};

struct RegressPtrs {
  float* soft;
  int64_t* amin;
  int64_t* amax;
  float* lse;
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

// ---- fp32 operands (kind::tf32).  MN-major 32-bit operands exist only in the SWIZZLE_128B_BASE32B layout:
// an atom = 4 channels (K rows, 128 B apart) x 32 pixels (128 contiguous bytes), whose 32-byte chunks are
// XOR-ed with the row number (address bits [5,7) ^= bits [7,9)); atoms of one 4-channel group follow each
// other along x (LBO = 512 B), groups follow each other at SBO = (T/32) * 512 B.  A warp stages one atom
// per step: lane -> (row = channel, 16-byte chunk), i.e. 4 x 128 contiguous global bytes and 512
// contiguous (permuted) shared bytes.  Returns through dst the same bytes the split pass revisits.
struct Atom32 {
  int kg, xa;   // 4-channel group, 32-pixel atom along x
  __device__ __forceinline__ void step(int nxa) {
    xa += 4;
    while (xa >= nxa) { xa -= nxa; ++kg; }
  }
};
__device__ __forceinline__ uint32_t atom32_off(int kg, int xa, int nxa, int cl, int q) {
  return (uint32_t)((kg * nxa + xa) * 512 + cl * 128 + ((((q >> 1) ^ cl) & 3) << 5) + ((q & 1) << 4));
}
__device__ __forceinline__ void stage_operand32(const FeatView& F, int64_t n, int y, int c0, int nch, int xs, int nxa, int W,
                                                unsigned char* dst, bool fast, int lt) {
  const int q = lt & 7, cl = (lt >> 3) & 3;
  const float* __restrict__ base =
      reinterpret_cast<const float*>(F.data) + n * F.sn + (int64_t)y * F.sh + (int64_t)(c0 + cl) * F.sc;
  const int nkg = nch >> 2;
  Atom32 a{0, lt >> 5};
  while (a.xa >= nxa) { a.xa -= nxa; ++a.kg; }
  for (; a.kg < nkg; a.step(nxa)) {
    const int x = xs + 32 * a.xa + 4 * q;
    const bool inside = fast && x >= 0 && x + 4 <= W;
    const bool empty = x + 4 <= 0 || x >= W;
    unsigned char* d = dst + atom32_off(a.kg, a.xa, nxa, cl, q);
    const float* src = base + (int64_t)(4 * a.kg) * F.sc;
    if (inside || empty)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(d)), "l"(inside ? src + x : src),
                   "r"(inside ? 16 : 0)
                   : "memory");
    else
      *reinterpret_cast<uint4*>(d) = load_chunk_slow<float>(src, x, W, F.sw);
  }
}
// split the chunks this thread staged into hi (in place) and lo (at +lo_off, a multiple of 1024 B: same
// swizzle phase).  hi keeps the top 11 significand bits, so it is exact in TF32 and lo = x - hi is exact in
// fp32; non-finite values keep hi = x (NaN canonicalised), lo = 0.
__device__ __forceinline__ void split_operand32(unsigned char* dst, int nch, int nxa, int lo_off, int lt) {
  const int q = lt & 7, cl = (lt >> 3) & 3;
  const int nkg = nch >> 2;
  Atom32 a{0, lt >> 5};
  while (a.xa >= nxa) { a.xa -= nxa; ++a.kg; }
  for (; a.kg < nkg; a.step(nxa)) {
    unsigned char* p = dst + atom32_off(a.kg, a.xa, nxa, cl, q);
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float f = __uint_as_float(w[i]);
      hi[i] = w[i] & 0xffffe000u;
      lo[i] = __float_as_uint(f - __uint_as_float(hi[i]));
      if (!(fabsf(f) < INFINITY)) { hi[i] = f != f ? 0x7fc00000u : w[i]; lo[i] = 0u; }
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
  if (g.d_fastest) {
    const uint32_t dchunk = t % (uint32_t)g.dchunks; t /= (uint32_t)g.dchunks;
    const uint32_t xt = t % (uint32_t)g.xtiles, row = t / (uint32_t)g.xtiles;
    const uint32_t n = row / (uint32_t)g.H;
    c.n = n; c.y = (int)(row - n * (uint32_t)g.H); c.xt = (int)xt; c.x0 = (int)xt * TC_TM; c.dc0 = (int)dchunk * g.dch;
    return c;
  }
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
  if (g.d_fastest) {
    dc0 += g.dch;
    if (dc0 < g.D) return;
    dc0 = 0;
    if (++xt < g.xtiles) { x0 += TC_TM; return; }
    xt = 0; x0 = 0;
    if (++y < g.H) return;
    y = 0; ++n;
    return;
  }
  if (++xt < g.xtiles) { x0 += TC_TM; return; }
  xt = 0; x0 = 0;
  if (++y < g.H) return;
  y = 0;
  if (++n < g.rows / g.H) return;
  n = 0;
  dc0 += g.dch;
}

template <typename Tin, typename Tout, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
inner_tc_kernel(FeatView L, FeatView R, Tout* __restrict__ out, RegressPtrs rp, TcGeom g, int fast,
                const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // swizzled atoms (TMA destinations, UMMA descriptors with base_offset 0) need 1024-byte alignment: 1 KB of slack
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr bool F32 = sizeof(Tin) == 4;
  constexpr int ES = (int)sizeof(Tin), EPC = 16 / ES;       // element bytes, elements per 16-byte chunk
  constexpr int KC = F32 ? TC_KC32 : TC_KC;                 // channels per stage
  unsigned char* stage0 = smem_raw;   // NSTAGE x { A: KC*128*ES | B: KC*ncol*ES } (fp32: hi copies, then lo copies)
  float* skew = reinterpret_cast<float*>(smem_raw + (g.nstage + (g.tma32 ? 2 : 0)) * (size_t)g.stage_bytes);   // epilogue scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(skew) + g.epi_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_NSTAGE + 6);
  const uint32_t smem_empty = smem_u32(bars), smem_full = smem_u32(bars + TC_NSTAGE),
                 tmem_full = smem_u32(bars + 2 * TC_NSTAGE), tmem_empty = smem_u32(bars + 2 * TC_NSTAGE + 2),
                 raw_full = smem_u32(bars + 2 * TC_NSTAGE + 4), lo_empty = smem_u32(bars + 3 * TC_NSTAGE + 4);

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
      mbar_init(lo_empty + 8 * i, 1);          // one UMMA commit
    }
    for (int i = 0; i < TC_NSTAGE; ++i) mbar_init(raw_full + 8 * i, 1);   // fp32 TMA: producer's expect_tx
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int nk = g.tma ? (g.C + g.boxc - 1) / g.boxc : (g.C + KC - 1) / KC;
  const int lo_off = KC * (TC_TM + g.ncol) * ES;            // fp32 only: hi -> lo distance inside a stage
  // contiguous tile range of this CTA (neighbouring x tiles share most of their right window in L2)
  int64_t per = (g.tiles + gridDim.x - 1) / gridDim.x;
  if (g.d_fastest) per = (per + g.dchunks - 1) / g.dchunks * g.dchunks;   // never split the chunks of one pixel tile
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
          if (F32 && g.tma32) {
            // boxes of 32 pixels x 16 channels as written by TMA (SWIZZLE_128B_ATOM_32B = the UMMA BASE32B
            // atoms): next 4 channels +512 B (SBO), next 32 pixels one 2 KB box further (LBO); hi operands in
            // ring stage s, lo operands in slot it & 1
            const uint32_t lo = smem_u32(stage0 + (size_t)nst * g.stage_bytes + (size_t)(it & 1) * g.stage_bytes);
            const uint32_t bofs = KC * TC_TM * ES;
            for (int ks = 0; ks < KC / 8; ++ks) {
              const uint64_t ahi = umma_desc(sA + ks * 1024, 2048, 512, 1), alo = umma_desc(lo + ks * 1024, 2048, 512, 1);
              const uint64_t bhi = umma_desc(sA + bofs + ks * 1024, 2048, 512, 1), blo = umma_desc(lo + bofs + ks * 1024, 2048, 512, 1);
              umma_tf32(td, alo, bhi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);   // small terms first
              umma_tf32(td, ahi, blo, idesc, 1u);
              umma_tf32(td, ahi, bhi, idesc, 1u);
            }
            umma_commit(lo_empty + 8 * (it & 1));                         // lo slot reusable once these complete
          } else if constexpr (F32) {
            // SWIZZLE_128B_BASE32B: LBO = next 32-pixel atom (512 B), SBO = next 4-channel group
            const uint32_t sboA32 = (TC_TM / 32) * 512, sboB32 = (uint32_t)(g.ncol / 32) * 512;
            for (int ks = 0; ks < nch / 8; ++ks) {                        // K = 8 per UMMA: two 4-channel groups
              const uint64_t ahi = umma_desc(sA + ks * 2 * sboA32, 512, sboA32, 1), alo = umma_desc(sA + lo_off + ks * 2 * sboA32, 512, sboA32, 1);
              const uint64_t bhi = umma_desc(sB + ks * 2 * sboB32, 512, sboB32, 1), blo = umma_desc(sB + lo_off + ks * 2 * sboB32, 512, sboB32, 1);
              umma_tf32(td, alo, bhi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);   // small terms first
              umma_tf32(td, ahi, blo, idesc, 1u);
              umma_tf32(td, ahi, bhi, idesc, 1u);
            }
          } else if (g.tma) {
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
    } else if (lane == 0 && g.tma32) {
      // fp32: raw k-chunks (10 box loads of 32 pixels x 16 channels) into the hi ring
      const uint32_t nst = (uint32_t)g.nstage, hb = (uint32_t)g.stage_bytes;
      uint32_t it = 0;
      TileCoord tc = tile_coord(t_beg, g);
      for (int64_t t = t_beg; t < t_end; ++t, tc.advance(g)) {
        const int xr0 = tc.x0 - tc.dc0 - g.dch;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const uint32_t s = it % nst, bar = raw_full + 8 * s;
          mbar_wait(smem_empty + 8 * s, ((it / nst) & 1) ^ 1);       // UMMAs that read this hi stage are done
          mbar_expect_tx(bar, hb);
          const uint32_t dstA = smem_u32(stage0 + (size_t)s * hb), dstB = dstA + KC * TC_TM * ES;
          const int c0 = kc * KC;
          for (int m = 0; m < TC_TM / 32; ++m) tma_load_4d(dstA + m * 2048, &tmL, bar, tc.x0 + 32 * m, tc.y, c0, (int)tc.n);
          for (int m = 0; m < g.ncol / 32; ++m) tma_load_4d(dstB + m * 2048, &tmR, bar, xr0 + 32 * m, tc.y, c0, (int)tc.n);
        }
      }
    }
  } else if (F32 && warp >= TC_EPI_WARPS && g.tma32) {
    // =============================================================== fp32: hi/lo splitter warps
    // All 128 threads split a landed raw chunk 16 bytes at a time -- hi in place, lo into slot job & 1 (free
    // once the UMMAs of job-2 have completed) -- and hand it to the issuer.
    const int lt = threadIdx.x - 32 * TC_EPI_WARPS;
    const uint32_t nst = (uint32_t)g.nstage;
    const uint32_t hb = (uint32_t)g.stage_bytes;
    const uint32_t njobs = (uint32_t)((t_end - t_beg) * nk);
    unsigned char* lo_base = stage0 + (size_t)nst * hb;
    for (uint32_t it = 0; it < njobs; ++it) {
      const uint32_t s = it % nst;
      unsigned char* hi = stage0 + (size_t)s * hb;
      unsigned char* lo = lo_base + (size_t)(it & 1) * hb;
      mbar_wait(raw_full + 8 * s, (it / nst) & 1);                       // the raw chunk has landed
      mbar_wait(lo_empty + 8 * (it & 1), ((it >> 1) & 1) ^ 1);           // UMMAs of job it-2 released the lo slot
      for (uint32_t o = 16u * lt; o < hb; o += 16u * 128u) {
        const uint4 v = *reinterpret_cast<const uint4*>(hi + o);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f = __uint_as_float(w[i]);
          h[i] = w[i] & 0xffffe000u;
          l[i] = __float_as_uint(f - __uint_as_float(h[i]));
          if (!(fabsf(f) < INFINITY)) { h[i] = f != f ? 0x7fc00000u : w[i]; l[i] = 0u; }
        }
        *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> async proxy (UMMA)
      mbar_arrive(smem_full + 8 * s);
    }
  } else if (warp >= TC_EPI_WARPS && (g.tma32 || (g.tma && g.nsplit == 2))) {
    // warps 8-11 have nothing to do (16-bit TMA path with the two-way epilogue)
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
      if constexpr (F32) {
        unsigned char* sA = stage0 + (size_t)(job % nst) * g.stage_bytes;
        const int kc = (int)(job % (uint32_t)nk), nch = min(KC, g.C - kc * KC);
        split_operand32(sA, nch, TC_TM / 32, lo_off, lt);
        split_operand32(sA + KC * TC_TM * ES, nch, g.ncol / 32, lo_off, lt);
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
        if constexpr (F32) {
          stage_operand32(L, tc.n, tc.y, c0, nch, tc.x0, TC_TM / 32, g.W, sA, fast, lt);
          stage_operand32(R, tc.n, tc.y, c0, nch, xr0, g.ncol / 32, g.W, sB, fast, lt);
        } else {
          stage_operand<Tin>(L, tc.n, tc.y, c0, nch, tc.x0, TC_TM / EPC, g.W, sA, fast, lt);
          stage_operand<Tin>(R, tc.n, tc.y, c0, nch, xr0, g.ncol / EPC, g.W, sB, fast, lt);
        }
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
    // (halves; thirds when warps 8-11 join in: 16-bit TMA volume path).  Disparity dl of lane t sits at column
    // 32q + t + dch - dl; this warp's window starts at 32q + cs with cs = dch - eb[hh+1] and is ncw columns wide
    // (>= part width + 32, multiple of 16).
    const int q = warp & 3, hh = warp >> 2;
    const int plo = hh == 0 ? g.eb[0] : hh == 1 ? g.eb[1] : g.eb[2];     // (static indices: no local copy of g)
    const int phi = hh == 0 ? g.eb[1] : hh == 1 ? g.eb[2] : g.eb[3];
    const int cs = g.dch - phi, ncw = (phi - plo + 32 + 15) / 16 * 16;
    const float inv = 1.f / (float)g.C, cnt = (float)g.C;
    uint32_t use = 0;
    // EPI_REGRESS: running softmax / arg-extrema state of this lane's pixel, carried across the disparity
    // chunks of a tile (chunks of one pixel tile are consecutive: d-fastest tile order)
    float m = -INFINITY, s = 0.f, ws = 0.f, minv = INFINITY, maxv = -INFINITY;
    int mini = 0x7fffffff, maxi = 0x7fffffff, nani = 0x7fffffff;
    uint32_t combines = 0;
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

      if constexpr (EPI == EPI_VOLUME) {
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
      } else {
        // ---- fused regression over this warp's disparities in ascending order (first index wins ties)
        if (tc.dc0 == 0) {
          m = -INFINITY; s = 0.f; ws = 0.f; minv = INFINITY; maxv = -INFINITY;
          mini = 0x7fffffff; maxi = 0x7fffffff; nani = 0x7fffffff;
        }
        int d0 = dlo;
        // full chunks of 8, lean form (the epilogue warps are issue-bound: three share a scheduler).  Extrema by
        // FMNMX trees, their first index by equality selects, one running-extremum update per chunk; the exponent
        // weights are compile-time k on top of a per-chunk float base (no I2F); NaNs are looked for only when the
        // chunk's sum of exponentials is NaN.  torch semantics as below: first index wins ties, NaN wins.
        for (; d0 + 8 <= dhi; d0 += 8) {
          float v[8];
          const bool allin = __all_sync(0xffffffffu, dz >= d0 + 8);        // no x < d fill in this chunk (usual)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float f = divide ? rp0[-(d0 + k)] * mul / cnt : rp0[-(d0 + k)] * mul;
            if (!allin) f = d0 + k < dz ? f : 0.f;                          // fill takes part (F8)
            v[k] = f;
          }
          const float lo01 = fminf(v[0], v[1]), lo23 = fminf(v[2], v[3]), lo45 = fminf(v[4], v[5]), lo67 = fminf(v[6], v[7]);
          const float hi01 = fmaxf(v[0], v[1]), hi23 = fmaxf(v[2], v[3]), hi45 = fmaxf(v[4], v[5]), hi67 = fmaxf(v[6], v[7]);
          const float cmin = fminf(fminf(lo01, lo23), fminf(lo45, lo67));
          const float cmax = fmaxf(fmaxf(hi01, hi23), fmaxf(hi45, hi67));
          int imin = 7, imax = 7;
#pragma unroll
          for (int k = 6; k >= 0; --k) {                                     // first k attaining the extremum
            imin = v[k] == cmin ? k : imin;
            imax = v[k] == cmax ? k : imax;
          }
          if (cmin < minv) { minv = cmin; mini = tc.dc0 + d0 + imin; }
          if (cmax > maxv) { maxv = cmax; maxi = tc.dc0 + d0 + imax; }
          const float mn = fmaxf(m, cmax);
          const float mnl = mn * kLog2e;
          const float a = (m == -INFINITY) ? 0.f : fast_exp2(fmaf(m, kLog2e, -mnl));
          float e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) e[k] = fast_exp2(fmaf(v[k], kLog2e, -mnl));
          const float S = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          const float T = (fmaf(2.f, e[2], e[1]) + fmaf(3.f, e[3], 4.f * e[4])) +
                          (fmaf(5.f, e[5], 6.f * e[6]) + 7.f * e[7]);                   // sum_k k * e_k
          const float fb = (float)(tc.dc0 + d0);
          s = fmaf(s, a, S);
          ws = fmaf(ws, a, fmaf(fb, S, T));
          m = mn;
          if (S != S) {                                                      // a NaN in the chunk (rare): first one
#pragma unroll
            for (int k = 7; k >= 0; --k)
              if (v[k] != v[k]) nani = min(nani, tc.dc0 + d0 + k);
          }
        }
        for (; d0 < dhi; d0 += 8) {                                          // ragged tail: element-wise form
          float v[8];
          float gm = -INFINITY;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int dl = d0 + k;
            float f = -INFINITY;
            if (dl < dhi) {
              f = dl < dz ? (divide ? rp0[-dl] * mul / cnt : rp0[-dl] * mul) : 0.f;   // fill takes part (F8)
              if (f < minv) { minv = f; mini = tc.dc0 + dl; }
              if (f > maxv) { maxv = f; maxi = tc.dc0 + dl; }
              if (f != f) nani = min(nani, tc.dc0 + dl);
            }
            v[k] = f;
            gm = fmaxf(gm, f);
          }
          const float mn = fmaxf(m, gm);
          const float mnl = mn * kLog2e;
          const float a = (m == -INFINITY) ? 0.f : fast_exp2(fmaf(m, kLog2e, -mnl));
          s *= a; ws *= a;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float e = fast_exp2(fmaf(v[k], kLog2e, -mnl));       // past the end: exp2(-inf) = 0
            s += e;
            ws = fmaf((float)(tc.dc0 + d0 + k), e, ws);
          }
          m = mn;
        }
        // after the last chunk, combine the warps of a quadrant: parts 1.. park their state, part 0 merges them
        // in ascending disparity order and stores (equal values: the smaller index wins, as torch does)
        if (tc.dc0 + g.dch < g.D) { __syncwarp(); continue; }
        float* pbase = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(skew) + g.epi_bytes - 2 * 2 * 128 * 8 * 4) +
                       (size_t)(combines++ & 1) * (2 * 128 * 8);
        if (hh >= 1) {
          float* part = pbase + ((size_t)(hh - 1) * 128 + 32 * q + lane) * 8;
          part[0] = m; part[1] = s; part[2] = ws; part[3] = minv; part[4] = maxv;
          part[5] = __int_as_float(mini); part[6] = __int_as_float(maxi); part[7] = __int_as_float(nani);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "r"(32 * g.nsplit) : "memory");
        if (hh == 0 && x < g.W) {
          for (int k = 0; k + 1 < g.nsplit; ++k) {
            const float* part = pbase + ((size_t)k * 128 + 32 * q + lane) * 8;
            const float m2 = part[0], s2 = part[1], w2 = part[2], minv2 = part[3], maxv2 = part[4];
            const int mini2 = __float_as_int(part[5]), maxi2 = __float_as_int(part[6]), nani2 = __float_as_int(part[7]);
            const float M = fmaxf(m, m2);
            const float a1 = (m == -INFINITY) ? 0.f : fast_exp2((m - M) * kLog2e);
            const float a2 = (m2 == -INFINITY) ? 0.f : fast_exp2((m2 - M) * kLog2e);
            s = s * a1 + s2 * a2; ws = ws * a1 + w2 * a2; m = M;
            if (minv2 < minv || (minv2 == minv && mini2 < mini)) { minv = minv2; mini = mini2; }
            if (maxv2 > maxv || (maxv2 == maxv && maxi2 < maxi)) { maxv = maxv2; maxi = maxi2; }
            nani = min(nani, nani2);
          }
          if (nani != 0x7fffffff) { mini = nani; maxi = nani; }
          const int64_t o = ((int64_t)tc.n * g.H + tc.y) * g.W + x;
          if (rp.soft) rp.soft[o] = ws / s;
          if (rp.lse) rp.lse[o] = m + __logf(s);
          if (rp.amin) rp.amin[o] = mini;
          if (rp.amax) rp.amax[o] = maxi;
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
  const int dq = fmt == 2 ? 32 : 16;      // fp32 operands come in 32-pixel atoms
  const int d16 = (int)((D + dq - 1) / dq * dq);
  g.dch = d16 < 128 ? d16 : 128;
  g.ncol = TC_TM + g.dch;
  g.xtiles = (int)ceil_div(W, TC_TM);
  g.dchunks = (int)ceil_div(D, g.dch);
  g.mean = mean;
  g.pow2 = (C & (C - 1)) == 0;
  g.fmt = fmt;
  g.tmem_buf = g.ncol <= 128 ? 128 : 256;
  g.stage_bytes = fmt == 2 ? 2 * TC_KC32 * (TC_TM + g.ncol) * 4 : TC_KC * (TC_TM + g.ncol) * 2;
  g.d_fastest = 0;
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
  const int es = fmt == 2 ? 4 : 2;
  const int64_t st[3] = {f.stride_h * es, f.stride_c * es, f.stride_n * es};   // bytes
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
  if (fmt == 2) {   // fp32: 32 pixels x 16 channels, 32-byte swizzle atoms (the UMMA SWIZZLE_128B_BASE32B layout)
    const cuuint32_t box[4] = {32, 1, (cuuint32_t)TC_KC32, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(f.data), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)g.boxc, 1};
  return enc(m, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(f.data),
             gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename Tin, typename Tout, int EPI>
static int launch_tc(const rsm_feat& left, const rsm_feat& right, void* out, RegressPtrs rp, const TcGeom& g_in,
                     cudaStream_t st, const char* where) {
  TcGeom g = g_in;
  // 16-bit operands by TMA when the views qualify (RSM_TC_TMA=0 keeps the cp.async loaders: A/B runs)
  alignas(64) CUtensorMap tmL, tmR;
  memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
  g.tma = 0; g.boxc = 0; g.nbb = 0; g.tma32 = 0;
  if (sizeof(Tin) == 4) {
    const char* e = getenv("RSM_TC_TMA");
    if (!(e && e[0] == '0') && make_tmap(&tmL, left, 2, g, g.rows / g.H) && make_tmap(&tmR, right, 2, g, g.rows / g.H)) {
      g.tma32 = 1;
      g.stage_bytes = TC_KC32 * (TC_TM + g.ncol) * 4;     // one raw/hi stage; two lo slots of the same size follow the ring
    }
  }
  if (sizeof(Tin) == 2) {
    const char* e = getenv("RSM_TC_TMA");
    g.boxc = g.C < TC_KC ? g.C : TC_KC;
    g.nbb = (g.ncol + 63) / 64;
    if (!(e && e[0] == '0') && make_tmap(&tmL, left, g.fmt, g, g.rows / g.H) && make_tmap(&tmR, right, g.fmt, g, g.rows / g.H)) {
      g.tma = 1;
      g.stage_bytes = (2 + g.nbb) * g.boxc * 128;
    }
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
    g.epi_bytes = 32 * 4 * nsplit * g.pitch * 4 + (EPI == EPI_REGRESS ? 2 * 2 * 128 * 8 * 4 : 0);   // rows (+ partials)
  };
  {
    const char* e = getenv("RSM_TC_NSPLIT");
    epilogue_parts(g.tma && !(e && e[0] == '2') ? 3 : 2);
    if (g.nsplit == 3 && 2 * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024 > 220 * 1024) epilogue_parts(2);
  }
  g.nstage = TC_NSTAGE;
  const int extra = g.tma32 ? 2 : 0;   // lo slots
  while (g.nstage > 2 && (g.nstage + extra) * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024 > 220 * 1024) --g.nstage;
  const size_t smem = (g.nstage + extra) * (size_t)g.stage_bytes + (size_t)g.epi_bytes + TC_BAR_BYTES + 1024;
  auto k = inner_tc_kernel<Tin, Tout, EPI>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return finish_launch(where);
  const int epc = 16 / (int)sizeof(Tin);
  const int fast = feat_vec16(left, epc) && feat_vec16(right, epc);   // 16-byte chunks start at multiples of epc elements
  const unsigned grid = (unsigned)(g.tiles < kNumSMs ? g.tiles : kNumSMs);   // persistent: one CTA per SM
  k<<<grid, TC_THREADS, smem, st>>>(view_of(left), view_of(right), (Tout*)out, rp, g, fast, tmL, tmR);
  return finish_launch(where);
}

static int tc_fmt(int in_dtype) { return in_dtype == RSM_F16 ? 0 : in_dtype == RSM_BF16 ? 1 : 2; }
// 16-bit: whole UMMAs of K = 16.  fp32 (3xTF32, K = 8) is opt-in with RSM_TC_FP32=1: parity-checked, but its
// operand pipeline (half the bytes in flight per stage, split pass in the loader warps) is still slower than
// the SIMT fp32 kernel -- see DESIGN.md section 4b.  The environment is read per call so tests can toggle it.
static bool tc_applies(int in_dtype, int64_t C, int64_t D) {
  if (C <= 0 || D <= 0) return false;
  if (in_dtype == RSM_F32) {
    const char* e = getenv("RSM_TC_FP32");
    return e && e[0] == '1' && C % 8 == 0;
  }
  return C % 16 == 0;
}

// returns RSM_ERR_UNSUPPORTED_CONFIG when the tensor-core path does not apply (caller falls back to SIMT)
int launch_inner_tc(const rsm_feat& left, const rsm_feat& right, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                    int64_t D, int mean, int in_dtype, int out_dtype, cudaStream_t st) {
  if (!tc_applies(in_dtype, C, D) || (in_dtype == RSM_F32 && out_dtype != RSM_F32)) return RSM_ERR_UNSUPPORTED_CONFIG;
  TcGeom g;
  if (int rc = tc_geom(N, C, H, W, D, mean, tc_fmt(in_dtype), g)) return rc;
  const RegressPtrs none{nullptr, nullptr, nullptr, nullptr};
  const char* where = "rsm_inner_fwd(tcgen05)";
  if (in_dtype == RSM_F32) return launch_tc<float, float, EPI_VOLUME>(left, right, out, none, g, st, where);
  if (in_dtype == RSM_F16) {
    if (out_dtype == RSM_F32) return launch_tc<__half, float, EPI_VOLUME>(left, right, out, none, g, st, where);
    return launch_tc<__half, __half, EPI_VOLUME>(left, right, out, none, g, st, where);
  }
  if (out_dtype == RSM_F32) return launch_tc<__nv_bfloat16, float, EPI_VOLUME>(left, right, out, none, g, st, where);
  return launch_tc<__nv_bfloat16, __nv_bfloat16, EPI_VOLUME>(left, right, out, none, g, st, where);
}

int launch_inner_regress_tc(const rsm_feat& left, const rsm_feat& right, int64_t N, int64_t C, int64_t H, int64_t W,
                            int64_t D, int mean, int in_dtype, const rsm_regress_out& out, cudaStream_t st) {
  if (!tc_applies(in_dtype, C, D)) return RSM_ERR_UNSUPPORTED_CONFIG;
  TcGeom g;
  if (int rc = tc_geom(N, C, H, W, D, mean, tc_fmt(in_dtype), g)) return rc;
  g.d_fastest = 1;   // the softmax state of a pixel lives in the epilogue's registers across its disparity chunks
  const RegressPtrs rp{(float*)out.soft, out.argmin, out.argmax, out.lse};
  const char* where = "rsm_inner_regress_fwd(tcgen05)";
  if (in_dtype == RSM_F32) return launch_tc<float, float, EPI_REGRESS>(left, right, nullptr, rp, g, st, where);
  if (in_dtype == RSM_F16) return launch_tc<__half, float, EPI_REGRESS>(left, right, nullptr, rp, g, st, where);
  return launch_tc<__nv_bfloat16, float, EPI_REGRESS>(left, right, nullptr, rp, g, st, where);
}

}  // namespace rsm
