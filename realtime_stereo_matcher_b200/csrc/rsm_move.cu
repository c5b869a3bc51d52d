// Data-movement volumes: concatenate, interweave, difference (forward + adjoint).
// All three are HBM-write bound: the kernels stage the tiny inputs on chip and
// emit coalesced 128-bit stores of the output in the reference's own layout.
#include <stdlib.h>

#include "rsm_common.cuh"

namespace rsm {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ T zero_of() { return from_f<T>(0.f); }

// ===================================================================== concatenate forward
// One CTA per output row (n, ch, y): the W source values go to shared memory once, then the
// row's W*D outputs (contiguous, D innermost) are written with VEC-wide stores.  The (x, d)
// decomposition of the running vector index is advanced incrementally (no division in the loop).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
concat_fwd_kernel(FeatView L, FeatView R, T* __restrict__ out, int C, int H, int W, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* srow = reinterpret_cast<T*>(smem_raw);

  const int64_t row = blockIdx.x;
  const int y = (int)(row % H);
  const int ch = (int)((row / H) % (2 * C));
  const int64_t n = row / ((int64_t)H * 2 * C);
  const bool right = ch >= C;
  const int c = right ? ch - C : ch;
  const FeatView& F = right ? R : L;
  const T* __restrict__ src = reinterpret_cast<const T*>(F.data) + n * F.sn + c * F.sc + y * F.sh;
  // the right row is kept REVERSED (srow[k] = R[W-1-k]) so that the D-run of a pixel, R[x-d] for
  // ascending d, is a run of ASCENDING shared-memory addresses starting at W-1-x+d0
  for (int x = threadIdx.x; x < W; x += kThreads) srow[right ? W - 1 - x : x] = __ldg(src + (int64_t)x * F.sw);
  __syncthreads();

  T* __restrict__ orow = out + row * (int64_t)W * D;
  const int DQ = D / VEC;
  const int nvec = W * DQ;
  const int xstep = kThreads / DQ, dstep = kThreads % DQ;
  int v = threadIdx.x;
  int x = v / DQ, dq = v - x * DQ;
  const T zero = zero_of<T>();
  for (; v < nvec; v += kThreads) {
    const int d0 = dq * VEC;
    if constexpr (VEC * sizeof(T) == 32) {
      // fp32, eight disparities per thread: one 256-bit streaming store (STG.E.ENL2.256)
      static_assert(sizeof(T) == 4, "256-bit path is fp32 only");
      float o[8];
      if (d0 + 7 <= x) {
        if (!right) {
          const float v0 = srow[x];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = v0;
        } else {
          const float* p = srow + (W - 1 - x + d0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = p[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (d0 + j <= x) ? srow[right ? W - 1 - x + d0 + j : x] : 0.f;
      }
      asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + (int64_t)v * 8), "f"(o[0]),
                   "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                   : "memory");
    } else if constexpr (VEC * sizeof(T) == 16) {
      Vec16<T> o;
      if (d0 + VEC - 1 <= x) {
        // interior vector (~all of them): no per-element range checks
        if (!right) {
          const T v0 = srow[x];
#pragma unroll
          for (int j = 0; j < VEC; ++j) o.v[j] = v0;
        } else {
          const int b = W - 1 - x + d0;
          if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) o.v[j] = srow[b + j];
          } else {
            // eight 16-bit values from an arbitrary 2-byte offset: five aligned words + funnel shifts
            const uint32_t* w32 = reinterpret_cast<const uint32_t*>(srow) + (b >> 1);
            const uint32_t sh = (uint32_t)(b & 1) * 16u;
            const uint32_t w0 = w32[0], w1 = w32[1], w2 = w32[2], w3 = w32[3], w4 = w32[4];
            o.raw = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                               __funnelshift_r(w3, w4, sh));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const int d = d0 + j;
          o.v[j] = (d <= x) ? srow[right ? W - 1 - x + d : x] : zero;
        }
      }
      stcs16(orow + (int64_t)v * VEC, o);
    } else {
      static_assert(VEC == 1, "scalar fallback only");
      __stcs(orow + v, (d0 <= x) ? srow[right ? W - 1 - x + d0 : x] : zero);
    }
    x += xstep;
    dq += dstep;
    if (dq >= DQ) { dq -= DQ; ++x; }
  }
}

// ===================================================================== concatenate adjoint
// gL[c,x] = sum_{d<=min(x,D-1)} gV[c,x,d];  gR[c,x'] = sum_{d<D, x'+d<W} gV[C+c,x'+d,d].
// One thread per (n,c,y,x): atomic-free, fp32 accumulation in ascending d.
template <typename T>
__global__ void __launch_bounds__(kThreads)
concat_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr, int64_t total,
                  int C, int H, int W, int D) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int c = (int)((i / ((int64_t)W * H)) % C);
  const int64_t n = i / ((int64_t)W * H * C);
  const int64_t rowL = ((n * 2 * C + c) * H + y) * (int64_t)W;        // in units of D
  const int64_t rowR = ((n * 2 * C + C + c) * H + y) * (int64_t)W;
  float sl = 0.f, sr = 0.f;
  const T* pl = gout + (rowL + x) * D;
  const int dl = min(x, D - 1);
  for (int d = 0; d <= dl; ++d) sl += to_f(__ldg(pl + d));
  const int dr = min(D - 1, W - 1 - x);
  const T* pr = gout + (rowR + x) * D;
  for (int d = 0; d <= dr; ++d) sr += to_f(__ldg(pr + (int64_t)d * (D + 1)));
  gl[i] = from_f<T>(sl);
  gr[i] = from_f<T>(sr);
}

// Row-tiled adjoint: one CTA per (n, channel of the 2C, y) row of the volume, i.e. a contiguous W x D
// matrix, streamed through shared memory TX pixels at a time with 4-byte LDGSTS (every load of the tile in
// flight, fully coalesced) into rows of an ODD word pitch, so that both reductions
//     left  half:  gL[x]  = sum_{d <= x}        M[x][d]        (row sums, masked like the forward fill)
//     right half:  gR[x'] = sum_{d, x'+d < W}   M[x'+d][d]     (diagonal sums; the tile carries a D-1 pixel halo)
// read shared memory conflict-free with one pixel per thread (two when the tile is wider than the CTA).  Sums run in
// ascending d (deterministic), D in chunks of <= 64 words per pixel.
// 16-bit volumes move as 32-bit words holding two disparities (D even).  (A variant with 16-byte copies,
// a pitch of 4 (mod 8) words and LDS.128 row sums measured 20% slower on B200: the reductions, not the
// copies, set the pace.)
template <typename T>
__device__ __forceinline__ float word_elem(uint32_t w, int d) {
  if constexpr (sizeof(T) == 4) return __uint_as_float(w);
  else {
    const unsigned short h = (d & 1) ? (unsigned short)(w >> 16) : (unsigned short)(w & 0xffffu);
    return to_f(*reinterpret_cast<const T*>(&h));
  }
}

// CWT = words per pixel and chunk when it is one of the usual ones (48: D = 48 / 96 / 192 in fp32, 96 / 192 in 16 bits;
// 24: D = 48 in 16 bits), else 0: with it the pitch is a constant and the full-length sums -- every pixel but the ones
// at the row's ends -- are unrolled with immediate offsets (the rolled loop spent ~6 instructions per element; ncu: 421 M
// warp instructions for 2.9 GB at cfg3).  RIGHT_ONLY: the grid covers the right half's rows only (the left half's plain
// row sums go through concat_bwd_left_kernel without shared memory).
template <typename T, bool CHUNKED, int CWT, bool RIGHT_ONLY>
__global__ void __launch_bounds__(kThreads)
concat_bwd_row_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr, int C, int H, int W,
                      int D, int TX, int P_, int DCH) {
  extern __shared__ __align__(16) uint32_t tile[];
  constexpr int EPW = 4 / (int)sizeof(T);          // elements per 32-bit word
  const int P = CWT ? (CWT | 1) : P_;
  const int DW = D / EPW;                          // words per pixel
  int64_t row = blockIdx.x;                        // (n, channel of the 2C, y)
  if (RIGHT_ONLY) {                                // blockIdx.x = (n, c, y) of the right half
    const int64_t nc = row / H;
    row = ((nc / C) * 2 * C + C + nc % C) * H + row % H;
  }
  const int y = (int)(row % H);
  const int ch = (int)((row / H) % (2 * C));
  const int64_t n = row / ((int64_t)H * 2 * C);
  const bool right = RIGHT_ONLY || ch >= C;
  const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(gout + row * (int64_t)W * D);
  T* __restrict__ dst = (right ? gr : gl) + ((n * C + (right ? ch - C : ch)) * H + y) * (int64_t)W;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  for (int x0 = 0; x0 < W; x0 += TX) {
    const int ntx = min(TX, W - x0);               // output pixels of this tile: threadIdx.x and threadIdx.x + kThreads
    float acc[2] = {0.f, 0.f};
    // disparity chunks of DCH (one chunk when D <= DCH): a chunk of the right half needs the pixels x' + d of ITS
    // disparities only, so the halo stays DCH - 1 pixels however large D is (a single pass over D = 192 re-read a
    // 191-pixel halo per 32-pixel tile: 7x the traffic)
    // (CHUNKED = false: one chunk = all of D, the staged words are contiguous in memory -- constants fold away)
    for (int d0 = 0; d0 < (CHUNKED ? D : 1); d0 += DCH) {
      const int dch = CHUNKED ? min(DCH, D - d0) : D, DWc = dch / EPW;
      const int start = x0 + (right ? d0 : 0);     // first pixel staged
      const int npx = min(ntx + (right ? dch - 1 : 0), W - start);
      // word w of the tile -> pixel w / DWc, word w % DWc of the chunk; a thread's words are kThreads apart, so the
      // shared-memory address, the source address and the wrap test advance by constants (no division in the loop)
      const int stepk = kThreads % DWc, stepp = kThreads / DWc;
      const uint32_t dstep = 4u * (stepp * P + stepk), wrapfix = 4u * (P - DWc);
      const int px0 = threadIdx.x / DWc, kk0 = threadIdx.x - px0 * DWc;
      __syncthreads();
      if (npx > 0) {
        int kk = kk0;
        uint32_t sdst = tile_s + 4u * (px0 * P + kk0);
        const uint32_t* g = src + (int64_t)(start + px0) * DW + d0 / EPW + kk0;
        const int64_t gstep = CHUNKED ? (int64_t)stepp * DW + stepk : kThreads;
        const int nword = npx * DWc;
        for (int w = threadIdx.x; w < nword; w += kThreads) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(g) : "memory");
          g += gstep; sdst += dstep; kk += stepk;
          if (kk >= DWc) {
            kk -= DWc; sdst += wrapfix;
            if (CHUNKED) g += DW - DWc;
          }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int xi = threadIdx.x + u * kThreads;
        if (xi >= ntx) break;
        const int x = x0 + xi;
        float a = acc[u];
        if (!right) {
          const uint32_t* r = tile + xi * P;
          const int nd = min(x - d0, dch - 1) + 1;             // disparities d0 .. of this chunk with d <= x
          if (CWT && nd == CWT * EPW) {
#pragma unroll
            for (int d = 0; d < CWT * EPW; ++d) a += word_elem<T>(r[d / EPW], d);
          } else {
            for (int d = 0; d < nd; ++d) a += word_elem<T>(r[d / EPW], d);
          }
        } else {
          const uint32_t* r = tile + xi * P;
          const int nd = min(dch, W - x - d0);                 // x + d0 + d < W
          if (CWT && nd == CWT * EPW) {
#pragma unroll
            for (int d = 0; d < CWT * EPW; ++d) a += word_elem<T>(r[d * (CWT | 1) + d / EPW], d);
          } else {
            for (int d = 0; d < nd; ++d) a += word_elem<T>(r[d * P + d / EPW], d);
          }
        }
        acc[u] = a;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int xi = threadIdx.x + u * kThreads;
      if (xi < ntx) dst[x0 + xi] = from_f<T>(acc[u]);
    }
  }
}

// Left half of the adjoint, gL[x] = sum_{d <= min(x, D-1)} gV[c, x, d]: plain sums of D contiguous elements per
// pixel -- no shared memory.  Four lanes per pixel take its 16-byte chunks in turn (a warp instruction reads eight
// pixels x 64 contiguous bytes: whole sectors), all of a lane's loads in flight, two shuffles fold the partial sums
// (fixed order: deterministic).  Needs D % (16 / sizeof(T)) == 0 and a 16-byte aligned gradient.
template <typename T>
__global__ void __launch_bounds__(kThreads)
concat_bwd_left_kernel(const T* __restrict__ gout, T* __restrict__ gl, int64_t npix, int C, int H, int W, int D) {
  constexpr int EPV = 16 / (int)sizeof(T);
  const int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 2;     // pixel (n, c, y, x) of gl
  const int q = threadIdx.x & 3;
  const bool live = i < npix;
  const int64_t ii = live ? i : npix - 1;
  const int64_t r = ii / W;                                                  // (n * C + c) * H + y
  const int x = (int)(ii - r * W);
  const int64_t ch = (int64_t)C * H;
  const int64_t n = r / ch;
  const T* __restrict__ src = gout + (((n * 2 * C) * H + (r - n * ch)) * W + x) * D;
  const int nchunk = D / EPV;
  float a = 0.f;
  constexpr int U = 4;                                                       // loads in flight per lane
  for (int c0 = q; c0 < nchunk; c0 += 4 * U) {
    Vec16<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c0 + 4 * u < nchunk) v[u] = ldcs16(src + (c0 + 4 * u) * EPV);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + 4 * u >= nchunk) break;
      const int d0 = (c0 + 4 * u) * EPV;
      if (x >= D - 1) {
#pragma unroll
        for (int e = 0; e < EPV; ++e) a += to_f(v[u].v[e]);
      } else {                                                               // the forward's fill region: d <= x only
#pragma unroll
        for (int e = 0; e < EPV; ++e) a += d0 + e <= x ? to_f(v[u].v[e]) : 0.f;
      }
    }
  }
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  if (live && q == 0) gl[i] = from_f<T>(a);
}

// Right half of the adjoint for rows that fit shared memory whole, gR[x'] = sum_{d, x'+d < W} gV[C+c, x'+d, d]:
// one CTA per (n, c, y) row, the contiguous W x D matrix fetched by ONE elected thread with bulk copies (TMA, completion
// counted on an mbarrier) -- no per-thread copy instructions, no halo, no padding; two or three CTAs per SM overlap
// one row's sums with the others' copies.  The diagonal of the DENSE matrix would put a warp's reads into two banks
// (pitch D = 16 mod 32 words); so lane l starts its diagonal at d = l and wraps around: word (x'+d) D + d with x' and d
// both advancing by one per lane is (2D+1) l + const -- an odd stride, conflict-free.  Sums are deterministic; their
// order starts at d = lane, not at 0.  (A persistent form -- one CTA per SM walking the rows through a ring of three
// row buffers -- measured slower: 690 vs 554 us for the whole adjoint at cfg3.)
template <typename T>
__global__ void __launch_bounds__(kThreads)
concat_bwd_right_bulk_kernel(const T* __restrict__ gout, T* __restrict__ gr, int C, int H, int W, int D, int TXP) {
  extern __shared__ __align__(128) unsigned char sraw[];
  // blockIdx.y = part of the row: output pixels [xa, xb), staged pixels [xa, min(W, xb + D - 1)) (contiguous in memory)
  const int xa = blockIdx.y * TXP, xb = min(W, xa + TXP), xe = min(W, xb + D - 1);
  const T* sM = reinterpret_cast<const T*>(sraw) - (int64_t)xa * D;   // [pixel][D], indexed by the row's pixel
  const uint32_t bytes = (uint32_t)(xe - xa) * (uint32_t)D * (uint32_t)sizeof(T);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sraw + bytes);
  const int64_t row = blockIdx.x;                          // (n * C + c) * H + y
  const int64_t nc = row / H;
  const int64_t grow = ((nc / C) * 2 * C + C + nc % C) * H + row % H;
  const unsigned char* __restrict__ src = reinterpret_cast<const unsigned char*>(gout + (grow * (int64_t)W + xa) * D);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sraw);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = min(32768u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + off), "l"(src + off), "r"(n), "r"(bar)
                   : "memory");
    }
  }
  {   // every thread waits for the bytes (bounded: a wait that expires traps instead of hanging)
    bool done = false;
    for (int it = 0; it < (1 << 12) && !done; ++it) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u), "r"(0x989680u) : "memory");
      done = ok != 0;
    }
    if (!done) __trap();
  }
  const int NT = blockDim.x;
  const int rot = (threadIdx.x & 31) % D;
  T* __restrict__ dst = gr + row * (int64_t)W;
  for (int x = xa + threadIdx.x; x < xb; x += NT) {
    const int nd = min(D, W - x);                           // x + d < W
    float a = 0.f;
    const T* p1 = sM + (int64_t)(x + rot) * D + rot;       // d = rot .. nd - 1
    const int e1 = nd - rot;
#pragma unroll 4
    for (int i = 0; i < e1; ++i) a += to_f(p1[i * (D + 1)]);
    const T* p2 = sM + (int64_t)x * D;                      // d = 0 .. min(rot, nd) - 1
    const int e2 = min(rot, nd);
#pragma unroll 4
    for (int i = 0; i < e2; ++i) a += to_f(p2[i * (D + 1)]);
    dst[x] = from_f<T>(a);
  }
}

// ============================================================================= interweave
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
interweave_fwd_kernel(FeatView L, FeatView R, T* __restrict__ out, int64_t total_vec, int C, int H,
                      int W) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total_vec) return;
  const int WV = W / VEC;
  const int xv = (int)(i % WV);
  const int y = (int)((i / WV) % H);
  const int ch = (int)((i / ((int64_t)WV * H)) % (2 * C));
  const int64_t n = i / ((int64_t)WV * H * 2 * C);
  const FeatView& F = (ch & 1) ? R : L;
  const T* src = reinterpret_cast<const T*>(F.data) + n * F.sn + (int64_t)(ch >> 1) * F.sc + y * F.sh;
  if constexpr (VEC > 1) {
    Vec16<T> v = ldcs16(src + (int64_t)xv * VEC);
    stcs16(out + i * VEC, v);
  } else {
    out[i] = __ldg(src + (int64_t)xv * F.sw);
  }
}

// 3-D grid version: blockIdx.y = output channel, blockIdx.z = batch item, blockIdx.x over the 16-byte vectors of one
// (H, W) plane, several per thread -- no 64-bit index arithmetic; planes of W-contiguous rows (also width-cropped
// views: one 32-bit division per vector for the row index)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
interweave_fwd_plane_kernel(FeatView L, FeatView R, T* __restrict__ out, int C, int H, int W, int dense) {
  const int ch = blockIdx.y;
  const int64_t n = blockIdx.z;
  const FeatView& F = (ch & 1) ? R : L;
  const T* __restrict__ src = reinterpret_cast<const T*>(F.data) + n * F.sn + (int64_t)(ch >> 1) * F.sc;
  T* __restrict__ dst = out + (n * 2 * C + ch) * (int64_t)H * W;
  const int WV = W / VEC, nvec = H * WV;
  constexpr int U = 4;
  const int v0 = (blockIdx.x * U) * kThreads + threadIdx.x;
  Vec16<T> t[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int v = v0 + u * kThreads;
    if (v < nvec) {
      if (dense) t[u] = ldcs16(src + (int64_t)v * VEC);
      else { const int y = v / WV, xv = v - y * WV; t[u] = ldcs16(src + (int64_t)y * F.sh + xv * VEC); }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int v = v0 + u * kThreads;
    if (v < nvec) stcs16(dst + (int64_t)v * VEC, t[u]);
  }
}

// adjoint with the same grid: plane (n, ch) of the gradient goes to gl (even ch) or gr (odd ch)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
interweave_bwd_plane_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr, int C, int nvec) {
  const int ch = blockIdx.y;
  const int64_t n = blockIdx.z;
  const T* __restrict__ src = gout + (n * 2 * C + ch) * (int64_t)nvec * VEC;
  T* __restrict__ dst = ((ch & 1) ? gr : gl) + (n * C + (ch >> 1)) * (int64_t)nvec * VEC;
  constexpr int U = 4;
  const int v0 = (blockIdx.x * U) * kThreads + threadIdx.x;
  Vec16<T> t[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (v0 + u * kThreads < nvec) t[u] = ldcs16(src + (int64_t)(v0 + u * kThreads) * VEC);
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (v0 + u * kThreads < nvec) stcs16(dst + (int64_t)(v0 + u * kThreads) * VEC, t[u]);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
interweave_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr,
                      int64_t total_vec, int C, int64_t plane_vec) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t p = i % plane_vec;                       // vector inside an (H,W) plane
  const int ch = (int)((i / plane_vec) % (2 * C));
  const int64_t n = i / (plane_vec * 2 * C);
  T* dst = ((ch & 1) ? gr : gl) + ((n * C + (ch >> 1)) * plane_vec + p) * VEC;
  if constexpr (VEC > 1) {
    stcs16(dst, ldcs16(gout + i * VEC));
  } else {
    *dst = gout[i];
  }
}

// ============================================================================= difference
// out (N,C,D,H,W): one thread per VEC consecutive x.  Inputs are D times smaller than the
// output and are served by L1/L2; the 128-bit output stores are the HBM stream.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
difference_fwd_kernel(FeatView L, FeatView R, T* __restrict__ out, int64_t total_vec, int C, int H,
                      int W, int D, float fill) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total_vec) return;
  const int WV = W / VEC;
  const int xv = (int)(i % WV);
  const int y = (int)((i / WV) % H);
  const int d = (int)((i / ((int64_t)WV * H)) % D);
  const int c = (int)((i / ((int64_t)WV * H * D)) % C);
  const int64_t n = i / ((int64_t)WV * H * D * C);
  const T* pl = reinterpret_cast<const T*>(L.data) + n * L.sn + c * L.sc + y * L.sh;
  const T* pr = reinterpret_cast<const T*>(R.data) + n * R.sn + c * R.sc + y * R.sh;
  const T fillv = from_f<T>(fill);
  T vals[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int x = xv * VEC + j;
    vals[j] = (x >= d) ? from_f<T>(to_f(__ldg(pl + (int64_t)x * L.sw)) - to_f(__ldg(pr + (int64_t)(x - d) * R.sw)))
                       : fillv;
  }
  if constexpr (VEC * sizeof(T) == 16) {
    Vec16<T> o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) o.v[j] = vals[j];
    stcs16(out + i * VEC, o);
  } else {
    out[i] = vals[0];
  }
}

// Row-block version of the difference volume: one CTA per (n, c, block of YB image rows).  The YB rows of both
// features are staged in shared memory once and every disparity plane's YB x W block -- contiguous in the
// (N,C,D,H,W) volume -- is written with 16-byte streaming stores, one warp per disparity.  No divisions on the
// store path, each input element is read from HBM once instead of D times through L1/L2.
// VEC consecutive pixels per store: 16 / 8 / 4 / 2 bytes, the widest one the row length and alignment allow
template <int BYTES> struct StoreWord;
template <> struct StoreWord<16> { using type = uint4; };
template <> struct StoreWord<8> { using type = uint2; };
template <> struct StoreWord<4> { using type = uint32_t; };
template <> struct StoreWord<2> { using type = unsigned short; };

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
difference_fwd_rows_kernel(FeatView L, FeatView R, T* __restrict__ out, int C, int H, int W, int D, float fill, int YB,
                           int yblocks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Word = typename StoreWord<VEC * (int)sizeof(T)>::type;
  T* sL = reinterpret_cast<T*>(smem_raw);
  constexpr int AL = 16 / (int)sizeof(T);
  const int padR = (D + 2 * AL - 1) / AL * AL;              // front margin: x - d < 0 stays inside the buffer
  T* sR = sL + ((size_t)YB * W + AL - 1) / AL * AL + padR;
  int64_t bid = blockIdx.x;
  const int yb = (int)(bid % yblocks); bid /= yblocks;
  const int c = (int)(bid % C);
  const int64_t n = bid / C;
  const int y0 = yb * YB, ny = min(YB, H - y0);
  const T* __restrict__ pl = reinterpret_cast<const T*>(L.data) + n * L.sn + (int64_t)c * L.sc + (int64_t)y0 * L.sh;
  const T* __restrict__ pr = reinterpret_cast<const T*>(R.data) + n * R.sn + (int64_t)c * R.sc + (int64_t)y0 * R.sh;
  for (int yy = threadIdx.x >> 5; yy < ny; yy += kThreads / 32)
    for (int x = threadIdx.x & 31; x < W; x += 32) {
      sL[yy * W + x] = __ldg(pl + (int64_t)yy * L.sh + (int64_t)x * L.sw);
      sR[yy * W + x] = __ldg(pr + (int64_t)yy * R.sh + (int64_t)x * R.sw);
    }
  __syncthreads();
  const int WV = W / VEC, nvec = ny * WV;
  const T fillv = from_f<T>(fill);
  const int lane = threadIdx.x & 31;
  for (int d = threadIdx.x >> 5; d < D; d += kThreads / 32) {
    T* __restrict__ o = out + ((((int64_t)n * C + c) * D + d) * H + y0) * (int64_t)W;
    int yy = lane / WV, xv = lane - yy * WV;                 // one division per disparity row, then incremental
    [[maybe_unused]] const int sh = (VEC - d % VEC) % VEC;    // (x - d) mod VEC for x % VEC == 0: warp-uniform
    for (int v = lane; v < nvec; v += 32) {
      const T* rl = sL + yy * W + xv * VEC;
      union { Word raw; T v[VEC]; } r;
      if constexpr (sizeof(T) == 4 && VEC == 4) {
        // right row shifted by d: two aligned 16-byte loads + a warp-uniform rotation instead of four strided
        // scalar loads (each of which would be a 4-way bank conflict)
        const Vec16<T> l = *reinterpret_cast<const Vec16<T>*>(rl);
        const T* ra = sR + yy * W + xv * VEC - d - sh;        // aligned quad holding x - d - sh .. (sh = 0: exact)
        const Vec16<T> a = *reinterpret_cast<const Vec16<T>*>(ra);
        const Vec16<T> b = *reinterpret_cast<const Vec16<T>*>(ra + VEC);
        float w[8] = {a.v[0], a.v[1], a.v[2], a.v[3], b.v[0], b.v[1], b.v[2], b.v[3]};
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float rv = sh == 0 ? w[j] : sh == 1 ? w[j + 1] : sh == 2 ? w[j + 2] : w[j + 3];
          r.v[j] = (xv * VEC + j >= d) ? l.v[j] - rv : fillv;
        }
      } else {
        const T* rr = sR + yy * W + xv * VEC - d;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          r.v[j] = (xv * VEC + j >= d) ? from_f<T>(to_f(rl[j]) - to_f(rr[j])) : fillv;
      }
      __stcs(reinterpret_cast<Word*>(o + (int64_t)v * VEC), r.raw);
      xv += 32;
      while (xv >= WV) { xv -= WV; ++yy; }
    }
  }
}

// gL[c,x] = sum_{d<=x} gV[c,d,x];  gR[c,x'] = -sum_{d, x'+d<W} gV[c,d,x'+d]
template <typename T>
__global__ void __launch_bounds__(kThreads)
difference_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr,
                      int64_t total, int C, int H, int W, int D) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int64_t nc = i / ((int64_t)W * H);
  const int64_t plane = (int64_t)H * W;
  const T* p = gout + nc * D * plane + (int64_t)y * W + x;
  float sl = 0.f, sr = 0.f;
  const int dl = min(x, D - 1);
  for (int d = 0; d <= dl; ++d) sl += to_f(__ldg(p + d * plane));
  const int dr = min(D - 1, W - 1 - x);
  for (int d = 0; d <= dr; ++d) sr -= to_f(__ldg(p + d * plane + d));
  gl[i] = from_f<T>(sl);
  gr[i] = from_f<T>(sr);
}

// ------------------------------------------------------------------------------- helpers
static bool feat_vec_ok(const rsm_feat& f, int vec, int esize) {
  return f.stride_w == 1 && f.stride_n % vec == 0 && f.stride_c % vec == 0 && f.stride_h % vec == 0 &&
         aligned_to(f.data, (size_t)vec * esize);
}
static bool grid_ok(int64_t blocks) { return blocks >= 0 && blocks <= 2147483647LL; }

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

extern "C" int rsm_concat_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C,
                              int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * C * H * W * D == 0) return RSM_OK;
  if (!left.data || !right.data || !out) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (W * D > 2147483647LL || !grid_ok(N * 2 * C * H)) return RSM_ERR_INVALID_SHAPE;
  if (!aligned_to(out, dtype_size(dtype))) return RSM_ERR_MISALIGNED;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const size_t smem = (size_t)W * sizeof(T) + 16;   // + one spare vector: the 16-bit path reads a fifth word
    if (smem > 200 * 1024) return (int)RSM_ERR_UNSUPPORTED_CONFIG;
    const dim3 grid((unsigned)(N * 2 * C * H));
    if constexpr (sizeof(T) == 4) {
      // measured on B200: 256-bit stores win 1-5 % up to D = 96 and lose ~3 % at D = 192
      if (D % 8 == 0 && D <= 128 && aligned_to(out, 32)) {
        auto k = concat_fwd_kernel<T, 8>;      // fp32: 256-bit stores
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, kThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)C, (int)H, (int)W, (int)D);
        return finish_launch("rsm_concat_fwd");
      }
    }
    if (D % VEC == 0 && aligned_to(out, 16)) {
      auto k = concat_fwd_kernel<T, VEC>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, kThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)C, (int)H, (int)W, (int)D);
    } else {
      auto k = concat_fwd_kernel<T, 1>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, kThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)C, (int)H, (int)W, (int)D);
    }
    return finish_launch("rsm_concat_fwd");
  });
}

extern "C" int rsm_concat_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                              int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  const int64_t total = N * C * H * W;
  if (total == 0) return RSM_OK;
  if (!gleft || !gright || (D > 0 && !gout)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(ceil_div(total, kThreads))) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    // row-tiled kernel: whole words per pixel, word-aligned rows (else the per-element gather)
    constexpr int EPW = 4 / (int)sizeof(T);
    if (D > 0 && D % EPW == 0 && aligned_to(gout, 4) && grid_ok(N * 2 * C * H)) {
      const int DW = (int)(D / EPW);
      const int CW = DW <= 64 ? DW : 48;                             // words per pixel and disparity chunk
      const int dch = CW * EPW;
      const int P = CW | 1;                                          // odd word pitch
      int64_t tx = (96 * 1024) / (4 * (int64_t)P) - (dch - 1);       // pixels per tile within 96 KB, halo included (48 KB tiles, five CTAs per SM: the same time)
      tx = tx < 32 ? 32 : tx;
      if (tx > 2 * kThreads) tx = 2 * kThreads;                      // two pixels per thread
      // equal tiles: W = 480 as 2 x 240 measured 0.65 of HBM against 0.61 for 454 + 26; a row that fits is one tile
      tx = ceil_div(W, ceil_div(W, tx));
      const size_t smem = (size_t)(tx + dch - 1) * P * 4;
      if (smem <= 200 * 1024) {
        // left half: streaming row sums when the chunks are 16-byte aligned, else through the row kernel as well
        constexpr int EPV = 16 / (int)sizeof(T);
        // (pixel rows of >= 192 bytes: at 96 bytes -- D = 48 in 16 bits -- the lanes' chunks do not divide evenly and the
        // split measured 6-10 % slower than one pass of the row kernel.  A skewed-scatter form of the RIGHT half -- rows read
        // straight from global memory, element d stored to sS[d][x - d], column sums -- measured slower than the row kernel
        // too: 447 vs 390 us at cfg3.)
        const bool split = D % EPV == 0 && D * (int64_t)sizeof(T) >= 192 && aligned_to(gout, 16) &&
                           grid_ok(ceil_div(4 * total, kThreads));
        if (split) {
          concat_bwd_left_kernel<T><<<(unsigned)ceil_div(4 * total, kThreads), kThreads, 0, st>>>(
              (const T*)gout, (T*)gleft, total, (int)C, (int)H, (int)W, (int)D);
          if (int rc = finish_launch("rsm_concat_bwd(left)")) return rc;
        }
        const int64_t row_bytes = W * D * (int64_t)sizeof(T);
        if (split && sizeof(T) == 4 && row_bytes % 16 == 0 && row_bytes <= 100 * 1024) {
          // right half, fp32: whole rows by bulk copy (two or more CTAs per SM), the row in equal passes of <= 256 threads
          // (16-bit rows: the same time as the row kernel, which keeps the ascending order of the sums)
          // a row in parts of <= 40 KB with their D-1 pixel halo (re-read from L2): five CTAs per SM instead of three
          const int64_t pix_bytes = D * (int64_t)sizeof(T);
          int64_t parts = ceil_div(row_bytes, 40 * 1024);
          if (parts > 1 && (W % parts != 0 || ((W / parts) * pix_bytes) % 16 != 0 || parts > 65535)) parts = 1;
          const int64_t txp = W / parts;
          const int64_t npx = txp + (parts > 1 ? D - 1 : 0) < W ? txp + (parts > 1 ? D - 1 : 0) : W;
          const int64_t passes = ceil_div(txp, kThreads);
          const unsigned nthr = (unsigned)(ceil_div(ceil_div(txp, passes), 32) * 32);
          const size_t smb = (size_t)(npx * pix_bytes) + 16;
          auto kb = concat_bwd_right_bulk_kernel<T>;
          if (smb > 48 * 1024) cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
          kb<<<dim3((unsigned)(N * C * H), (unsigned)parts), nthr, smb, st>>>((const T*)gout, (T*)gright, (int)C, (int)H, (int)W,
                                                                             (int)D, (int)txp);
          return finish_launch("rsm_concat_bwd(right, bulk)");
        }
        const unsigned rows = (unsigned)(N * (split ? 1 : 2) * C * H);
        auto go = [&](auto k) -> int {
          if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          k<<<rows, kThreads, smem, st>>>((const T*)gout, (T*)gleft, (T*)gright, (int)C, (int)H, (int)W, (int)D, (int)tx, P, dch);
          return finish_launch("rsm_concat_bwd");
        };
        const bool chunked = CW != DW;
        if (split) {
          if (CW == 48) return chunked ? go(concat_bwd_row_kernel<T, true, 48, true>) : go(concat_bwd_row_kernel<T, false, 48, true>);
          if (CW == 24) return go(concat_bwd_row_kernel<T, false, 24, true>);
          return chunked ? go(concat_bwd_row_kernel<T, true, 0, true>) : go(concat_bwd_row_kernel<T, false, 0, true>);
        }
        if (CW == 48) return chunked ? go(concat_bwd_row_kernel<T, true, 48, false>) : go(concat_bwd_row_kernel<T, false, 48, false>);
        if (CW == 24) return go(concat_bwd_row_kernel<T, false, 24, false>);
        return chunked ? go(concat_bwd_row_kernel<T, true, 0, false>) : go(concat_bwd_row_kernel<T, false, 0, false>);
      }
    }
    concat_bwd_kernel<T><<<(unsigned)ceil_div(total, kThreads), kThreads, 0, st>>>(
        (const T*)gout, (T*)gleft, (T*)gright, total, (int)C, (int)H, (int)W, (int)D);
    return finish_launch("rsm_concat_bwd");
  });
}

extern "C" int rsm_interweave_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C,
                                  int64_t H, int64_t W, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * C * H * W == 0) return RSM_OK;
  if (!left.data || !right.data || !out) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const bool vec = W % VEC == 0 && feat_vec_ok(left, VEC, sizeof(T)) && feat_vec_ok(right, VEC, sizeof(T)) &&
                     aligned_to(out, 16);
    if (vec && 2 * C <= 65535 && N <= 65535 && H * (W / VEC) < (1LL << 30)) {
      const int nvec = (int)(H * (W / VEC));
      const dim3 grid((unsigned)ceil_div(nvec, 4 * kThreads), (unsigned)(2 * C), (unsigned)N);
      const int dense = left.stride_h == W && right.stride_h == W;
      interweave_fwd_plane_kernel<T, VEC><<<grid, kThreads, 0, st>>>(view_of(left), view_of(right), (T*)out, (int)C, (int)H,
                                                                   (int)W, dense);
      return finish_launch("rsm_interweave_fwd");
    }
    const int64_t total = N * 2 * C * H * (vec ? W / VEC : W);
    if (!grid_ok(ceil_div(total, kThreads))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, kThreads);
    if (vec)
      interweave_fwd_kernel<T, VEC><<<blocks, kThreads, 0, st>>>(view_of(left), view_of(right), (T*)out, total,
                                                                 (int)C, (int)H, (int)W);
    else
      interweave_fwd_kernel<T, 1><<<blocks, kThreads, 0, st>>>(view_of(left), view_of(right), (T*)out, total,
                                                               (int)C, (int)H, (int)W);
    return finish_launch("rsm_interweave_fwd");
  });
}

extern "C" int rsm_interweave_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                                  int64_t H, int64_t W, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * C * H * W == 0) return RSM_OK;
  if (!gout || !gleft || !gright) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t plane = H * W;
    const bool vec = plane % VEC == 0 && aligned_to(gout, 16) && aligned_to(gleft, 16) && aligned_to(gright, 16);
    const int64_t plane_vec = vec ? plane / VEC : plane;
    if (vec && 2 * C <= 65535 && N <= 65535 && plane_vec < (1LL << 30)) {
      const dim3 grid((unsigned)ceil_div(plane_vec, 4 * kThreads), (unsigned)(2 * C), (unsigned)N);
      interweave_bwd_plane_kernel<T, VEC><<<grid, kThreads, 0, st>>>((const T*)gout, (T*)gleft, (T*)gright, (int)C, (int)plane_vec);
      return finish_launch("rsm_interweave_bwd");
    }
    const int64_t total = N * 2 * C * plane_vec;
    if (!grid_ok(ceil_div(total, kThreads))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, kThreads);
    if (vec)
      interweave_bwd_kernel<T, VEC><<<blocks, kThreads, 0, st>>>((const T*)gout, (T*)gleft, (T*)gright, total,
                                                                 (int)C, plane_vec);
    else
      interweave_bwd_kernel<T, 1><<<blocks, kThreads, 0, st>>>((const T*)gout, (T*)gleft, (T*)gright, total,
                                                               (int)C, plane_vec);
    return finish_launch("rsm_interweave_bwd");
  });
}

extern "C" int rsm_difference_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C,
                                  int64_t H, int64_t W, int64_t D, float fill, int dtype, int device,
                                  void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * C * H * W * D == 0) return RSM_OK;
  if (!left.data || !right.data || !out) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const bool vec = W % VEC == 0 && aligned_to(out, 16);
    // row-block kernel: 8 image rows per CTA, the widest store the row length and the output alignment allow
    // (rows too long for its shared-memory tile take the per-vector kernel)
    {
      const int YB = 8;
      constexpr int AL = 16 / (int)sizeof(T);
      const int64_t yblocks = ceil_div(H, YB), bx = N * C * yblocks;
      const size_t smem = ((size_t)2 * (YB * W + AL) + (D + 2 * AL - 1) / AL * AL + AL) * sizeof(T);
      if (smem <= 96 * 1024 && grid_ok(bx)) {
        auto launch = [&](auto k) -> int {
          if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          k<<<(unsigned)bx, kThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)C, (int)H, (int)W, (int)D, fill,
                                                  YB, (int)yblocks);
          return finish_launch("rsm_difference_fwd");
        };
        auto fits = [&](int v) { return W % v == 0 && aligned_to(out, (size_t)v * sizeof(T)); };
        if (fits(AL)) return launch(difference_fwd_rows_kernel<T, AL>);
        if (AL >= 4 && fits(AL / 2)) return launch(difference_fwd_rows_kernel<T, AL / 2>);
        if (AL >= 8 && fits(AL / 4)) return launch(difference_fwd_rows_kernel<T, AL / 4>);
        return launch(difference_fwd_rows_kernel<T, 1>);
      }
    }
    const int64_t total = N * C * D * H * (vec ? W / VEC : W);
    if (!grid_ok(ceil_div(total, kThreads))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, kThreads);
    if (vec)
      difference_fwd_kernel<T, VEC><<<blocks, kThreads, 0, st>>>(view_of(left), view_of(right), (T*)out, total,
                                                                 (int)C, (int)H, (int)W, (int)D, fill);
    else
      difference_fwd_kernel<T, 1><<<blocks, kThreads, 0, st>>>(view_of(left), view_of(right), (T*)out, total,
                                                               (int)C, (int)H, (int)W, (int)D, fill);
    return finish_launch("rsm_difference_fwd");
  });
}

extern "C" int rsm_difference_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                                  int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  const int64_t total = N * C * H * W;
  if (total == 0) return RSM_OK;
  if (!gleft || !gright || (D > 0 && !gout)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(ceil_div(total, kThreads))) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    difference_bwd_kernel<T><<<(unsigned)ceil_div(total, kThreads), kThreads, 0, st>>>(
        (const T*)gout, (T*)gleft, (T*)gright, total, (int)C, (int)H, (int)W, (int)D);
    return finish_launch("rsm_difference_bwd");
  });
}
