// MobileStereoNetV4 head, fused: F.interpolate(cost[:,None], [D,H,W], 'trilinear') -> softmax over D
// -> expectation (model/mobile_stereo_net_v4.py:511-518), forward and adjoint, without ever
// materialising the (B,D,H,W) tensor.  The kernel is bound by instruction issue and MUFU (one
// exp per fine disparity), not by HBM, so the work per fine disparity is stripped
// to: one FMA (lerp along d of two pre-scaled coarse slices), one ex2, and the two accumulations.
//
// The interpolation is separable and is evaluated in that order per CTA (32x8 output pixels):
//   1. the coarse footprint of the tile is staged in shared memory (raw[k][fy][fx]);
//   2. it is interpolated along x ONCE for the tile's 32 output columns (rows[k][fy][tx]);
//   3. each thread (one pixel) gets slice value c_k with one y-lerp of two conflict-free LDS;
//   4. fine disparities are visited interval by interval: all d' whose source index i0(d') equals
//      k interpolate between c_k and c_{k+1}.  For the x4 head (D == 4*Dc) the intervals are
//      regular -- weights 1/8,3/8,5/8,7/8 -- and the loop body is fully unrolled with constant
//      weights; other ratios use per-CTA index/weight tables in shared memory.
#include <type_traits>

#include "rsm_common.cuh"

namespace rsm {

// Source index/weights of F.interpolate(align_corners=False) as ATen computes them in fp32
// (no FMA contraction, so the oracle's table matches bit for bit).
struct Lin {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lin lin_index(int o, float scale, int n_in) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)o, 0.5f)), 0.5f);
  src = src < 0.f ? 0.f : src;
  Lin r;
  r.i0 = min((int)src, n_in - 1);
  r.i1 = r.i0 + (r.i0 < n_in - 1 ? 1 : 0);
  r.w1 = fminf(fmaxf(__fsub_rn(src, (float)r.i0), 0.f), 1.f);
  r.w0 = __fsub_rn(1.f, r.w1);
  return r;
}

// lin_index(o, 0.25f, n_in) in integers -- the x4 head: src = (o - 1.5) / 4, weights k / 8, all exact in fp32, so the
// values are the same bit for bit (for o inside the image; rows past it are never stored)
__device__ __forceinline__ Lin lin_index_x4(int o, int n_in) {
  Lin r;
  const int t = o - 2;
  r.i0 = t < 0 ? 0 : min(t >> 2, n_in - 1);
  r.i1 = r.i0 + (r.i0 < n_in - 1 ? 1 : 0);
  r.w1 = t < 0 ? 0.f : 0.125f + 0.25f * (float)(t & 3);
  r.w0 = 1.f - r.w1;
  return r;
}

constexpr int kTX = 32, kTY = 8;   // fine-pixel tile of one CTA (256 threads, one pixel each; 32x16 measured slower)
constexpr int kNT = kTX * kTY;

struct TailGeom {
  int Dc, Hc, Wc, D, H, W;
  int FH, FW;        // coarse rows / cols a tile can touch (upper bound)
  int fast4;         // D == 4 * Dc
  int fastx;         // W == 4 * Wc
  int fasty;         // H == 4 * Hc
  int all4;          // x4 along all three axes with the full 4 x 10 coarse footprint: specialised staging + stabiliser
  float sd, sh, sw;  // in/out scale per axis
};

// shared memory: [ rows Dc*FH*32 | raw Dc*FH*FW | w1tab D | dstart Dc+1 | i0tab D ]  (tables: generic ratio)
//            or  [ rows Dc*4*32  | raw Dc*40 (+pad), then part 2*8*128 over it | rowmax 2*128 ]   (all-x4 path)
struct TailSmem {
  float* rows;
  float* raw;
  float* w1tab;
  int* dstart;
  int* i0tab;
  float* part;     // all4: per-warp column maxima [8][128], then minima [8][128], of rows
  float* rowmax;   // all4: max over k of rows[k][fy][tx] [128], then the min [128]
  __device__ __forceinline__ TailSmem(float* base, const TailGeom& g) {
    rows = base;
    raw = rows + g.Dc * g.FH * kTX;
    w1tab = raw + g.Dc * g.FH * g.FW;
    dstart = reinterpret_cast<int*>(w1tab + g.D);
    i0tab = dstart + g.Dc + 1;
    // all4: the per-warp extrema overlay the raw footprint (dead once the x-interpolation has read it: a barrier in
    // stage_tile_all4 separates the two) -- 33.8 instead of 41.5 KB at Dc = 48: six CTAs per SM instead of five
    part = raw;
    rowmax = raw + max((g.Dc * g.FH * g.FW + 3) & ~3, 2 * 8 * 4 * kTX);
  }
};
static size_t tail_smem_bytes(const TailGeom& g) {
  size_t n = (size_t)g.Dc * g.FH * (kTX + g.FW);
  if (g.all4) {
    const size_t raw = ((size_t)g.Dc * g.FH * g.FW + 3) & ~(size_t)3, part = 2 * 8 * 4 * kTX;
    n = (size_t)g.Dc * g.FH * kTX + (raw > part ? raw : part) + 2 * 4 * kTX;
  }
  else if (!g.fast4) n += (size_t)g.D + g.Dc + 1 + g.D;
  return n * sizeof(float);
}

// Steps 1 + 2 (+ tables).  Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void stage_tile(const T* __restrict__ cost_b, const TailSmem& sm, const TailGeom& g,
                                           int cy0, int cx0) {
  // ---- 1. raw footprint: thread -> (r = tid % 64, k = tid / 64 + 4j); no divisions in the loop
  const int per = g.FH * g.FW;
  const int64_t plane = (int64_t)g.Hc * g.Wc;
  for (int r = threadIdx.x & 63; r < per; r += 64) {
    const int fy = r / g.FW, fx = r - fy * g.FW;
    const int cy = min(cy0 + fy, g.Hc - 1), cx = min(cx0 + fx, g.Wc - 1);
    const T* __restrict__ src = cost_b + (int64_t)cy * g.Wc + cx;
    constexpr int U = 4, KS = kNT / 64;   // KS slices are staged concurrently by the 64-thread groups
    for (int k0 = threadIdx.x >> 6; k0 < g.Dc; k0 += KS * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (k0 + KS * u < g.Dc) ? to_f(__ldg(src + (k0 + KS * u) * plane)) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (k0 + KS * u < g.Dc) sm.raw[(k0 + KS * u) * per + r] = v[u];
    }
  }
  if (!g.fast4) {
    for (int d = threadIdx.x; d < g.D; d += kNT) {
      const Lin ld = lin_index(d, g.sd, g.Dc);
      sm.w1tab[d] = ld.w1;
      sm.i0tab[d] = ld.i0;
    }
  }
  __syncthreads();
  // ---- 2. interpolate along x for this tile's 32 columns: rows[(k*FH + fy)*32 + tx]
  if (g.fastx) {
    // W == 4 * Wc: four consecutive output columns 4j..4j+3 need only raw[j-1], raw[j], raw[j+1] (clamped)
    // with the constant weights of lin_index -- 3 LDS + one 128-bit store per 4 outputs
    const int q = threadIdx.x & 7;                                  // column quad of the tile
    const int j = (blockIdx.x * kTX) / 4 + q;                       // coarse column of outputs 4j+2, 4j+3
    const int jm = min(max(j - 1, 0), g.Wc - 1), jc = min(j, g.Wc - 1), jp = min(j + 1, g.Wc - 1);
    const int am = min(jm - cx0, g.FW - 1), ac = min(jc - cx0, g.FW - 1), ap = min(jp - cx0, g.FW - 1);
    const int npair = g.Dc * g.FH;
    for (int p = threadIdx.x >> 3; p < npair; p += kNT / 8) {
      const float* rr = sm.raw + p * g.FW;
      const float vm = rr[am], vc = rr[ac], vp = rr[ap];
      float4 o;
      o.x = 0.375f * vm + 0.625f * vc;
      o.y = 0.125f * vm + 0.875f * vc;
      o.z = 0.875f * vc + 0.125f * vp;
      o.w = 0.625f * vc + 0.375f * vp;
      *reinterpret_cast<float4*>(sm.rows + p * kTX + 4 * q) = o;
    }
  } else {
    const int tx = threadIdx.x & (kTX - 1);
    const Lin lx = lin_index(blockIdx.x * kTX + tx, g.sw, g.Wc);
    const int a0 = min(lx.i0 - cx0, g.FW - 1), a1 = min(lx.i1 - cx0, g.FW - 1);
    const int npair = g.Dc * g.FH;          // (k, fy) pairs; raw row of pair p starts at p * FW
    for (int p = threadIdx.x >> 5; p < npair; p += kNT / kTX) {
      const float* rr = sm.raw + p * g.FW;
      sm.rows[p * kTX + tx] = lx.w0 * rr[a0] + lx.w1 * rr[a1];
    }
  }
  if (!g.fast4) {
    for (int k = threadIdx.x; k <= g.Dc; k += kNT) {   // dstart[k] = #{d : i0(d) < k}
      int lo = 0, hi = g.D;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sm.i0tab[mid] < k) lo = mid + 1; else hi = mid;
      }
      sm.dstart[k] = lo;
    }
  }
  __syncthreads();
}

// Steps 1 + 2 for the x4 x4 x4 head (footprint exactly 4 x 10 coarse pixels per slice): division-free
// staging with every load of a batch in flight, and, on the way, rowmax[fy][tx] = max_k rows[k][fy][tx].
// A fine value is a convex combination of two rows entries, so max(rowmax[fy0], rowmax[fy1]) bounds every
// fine value of the pixel from above: the softmax stabiliser without a pass over the 48 slices.
template <typename T>
__device__ __forceinline__ void stage_tile_all4(const T* __restrict__ cost_b, const TailSmem& sm, const TailGeom& g,
                                                int cy0, int cx0) {
  constexpr int FW = 10;
  const int plane = g.Hc * g.Wc;           // Dc * Hc * Wc < 2^31 (host check): 32-bit offsets inside a batch item
  // ---- 1. raw[(k*4 + fy)*10 + fx]: a thread keeps its footprint cell (fy, fx) and walks the slices six apart (240 of
  // the 256 threads; one division per thread, not per element: the flat-index form spent 25 instructions per value)
  {
    const int kk = threadIdx.x / 40, r = threadIdx.x - kk * 40;
    if (kk < 6) {
      const int fy = r / FW, fx = r - fy * FW;
      const int cy = min(cy0 + fy, g.Hc - 1), cx = min(cx0 + fx, g.Wc - 1);
      const T* __restrict__ src = cost_b + ((int64_t)kk * plane + cy * g.Wc + cx);
      float* dst = sm.raw + threadIdx.x;                 // (kk * 4 + fy) * 10 + fx == kk * 40 + r
      const int64_t step = 6 * (int64_t)plane;
      int k = kk;
      for (; k + 18 < g.Dc; k += 24, src += 4 * step, dst += 4 * 240) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = to_f(__ldg(src + u * step));
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u * 240] = v[u];
      }
      for (; k < g.Dc; k += 6, src += step, dst += 240) *dst = to_f(__ldg(src));
    }
  }
  __syncthreads();
  // ---- 2. x-interpolation (constant weights, see stage_tile) + running column maxima
  const int q = threadIdx.x & 7, p0 = threadIdx.x >> 3;     // p0 = 4 * warp + fy
  const int j = (blockIdx.x * kTX) / 4 + q;
  const int jm = min(max(j - 1, 0), g.Wc - 1), jc = min(j, g.Wc - 1), jp = min(j + 1, g.Wc - 1);
  const int am = min(jm - cx0, FW - 1), ac = min(jc - cx0, FW - 1), ap = min(jp - cx0, FW - 1);
  const int npair = g.Dc * 4;
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
  const float* rr = sm.raw + p0 * FW;
  float* ro = sm.rows + p0 * kTX + 4 * q;
#pragma unroll 2
  for (int p = p0; p < npair; p += kNT / 8, rr += (kNT / 8) * FW, ro += (kNT / 8) * kTX) {
    const float vm = rr[am], vc = rr[ac], vp = rr[ap];
    float4 o;
    o.x = 0.375f * vm + 0.625f * vc;
    o.y = 0.125f * vm + 0.875f * vc;
    o.z = 0.875f * vc + 0.125f * vp;
    o.w = 0.625f * vc + 0.375f * vp;
    *reinterpret_cast<float4*>(ro) = o;
    mx.x = fmaxf(mx.x, o.x); mx.y = fmaxf(mx.y, o.y); mx.z = fmaxf(mx.z, o.z); mx.w = fmaxf(mx.w, o.w);
    mn.x = fminf(mn.x, o.x); mn.y = fminf(mn.y, o.y); mn.z = fminf(mn.z, o.z); mn.w = fminf(mn.w, o.w);
  }
  // (fmaxf / fminf skip NaNs: a NaN column keeps a finite range, and the NaN then propagates through pass 2)
  __syncthreads();   // every thread has read its share of raw: part overlays it
  {
    float* pp = sm.part + (p0 >> 2) * (4 * kTX) + (p0 & 3) * kTX + 4 * q;
    *reinterpret_cast<float4*>(pp) = mx;
    *reinterpret_cast<float4*>(pp + 8 * 4 * kTX) = mn;
  }
  __syncthreads();
  {   // threads 0-127: column maxima, threads 128-255: column minima
    const int c = threadIdx.x & (4 * kTX - 1);
    const bool lo = threadIdx.x >= 4 * kTX;
    const float* pp = sm.part + (lo ? 8 * 4 * kTX : 0) + c;
    float m = pp[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = lo ? fminf(m, pp[w * (4 * kTX)]) : fmaxf(m, pp[w * (4 * kTX)]);
    sm.rowmax[threadIdx.x] = m;
  }
  __syncthreads();
}

// Step 3: the y-lerp of one output pixel; slice k is at base + k * stride
struct SliceY {
  const float* p0;   // rows + (y0 - cy0) * 32 + tx
  const float* p1;   // rows + (y1 - cy0) * 32 + tx
  float wy0, wy1;
  int stride;        // FH * 32
  __device__ __forceinline__ SliceY(const TailSmem& sm, const TailGeom& g, int y, int cy0) {
    const Lin ly = lin_index(y, g.sh, g.Hc);
    const int tx = threadIdx.x & (kTX - 1);
    p0 = sm.rows + min(ly.i0 - cy0, g.FH - 1) * kTX + tx;
    p1 = sm.rows + min(ly.i1 - cy0, g.FH - 1) * kTX + tx;
    wy0 = ly.w0; wy1 = ly.w1;
    stride = g.FH * kTX;
  }
  __device__ __forceinline__ float get(int k) const { return wy0 * p0[k * stride] + wy1 * p1[k * stride]; }
  // all4: compile-time slice stride (4 rows x 32 columns) -> LDS with immediate offsets
  __device__ __forceinline__ float get4(int k) const { return wy0 * p0[k * (4 * kTX)] + wy1 * p1[k * (4 * kTX)]; }
  // c_k * log2(e) - M * log2(e)
  template <bool ALL4>
  __device__ __forceinline__ float scaled(int k, float neg_ml) const {
    if constexpr (ALL4) return fmaf(wy0 * kLog2e, p0[k * (4 * kTX)], fmaf(wy1 * kLog2e, p1[k * (4 * kTX)], neg_ml));
    else return fmaf(get(k), kLog2e, neg_ml);
  }
  __device__ __forceinline__ float bound(const TailSmem& sm) const {   // all4: upper bound of every fine value
    return fmaxf(sm.rowmax[p0 - sm.rows], sm.rowmax[p1 - sm.rows]);
  }
  __device__ __forceinline__ float lower_bound(const TailSmem& sm) const {
    return fminf(sm.rowmax[4 * kTX + (p0 - sm.rows)], sm.rowmax[4 * kTX + (p1 - sm.rows)]);
  }
};

struct NoTrack {
  __device__ __forceinline__ void update(float, int) {}
  __device__ __forceinline__ void update_min(float, int) {}
  __device__ __forceinline__ void update_max(float, int) {}
};

// ===================================================================================== forward
// Pass 2 of the x4 head: intervals in ascending d.  Slices are pre-scaled, cs = c*log2(e) - M*log2(e), so
// exp(f - M) = ex2(lerp(cs0, cs1)); the four fine values of an interval are equally spaced
// (f_j = cs0 + (2j+1)/8 * dl), so their exponentials form a geometric progression: two ex2 (MUFU is the
// scarcest pipe here) + three multiplies instead of four ex2.
//   ROBUST = false: ascending progression from e_0 with ratio 2^(dl/4).  Only valid when no f_j can flush,
//                   i.e. the caller has bounded the pixel's whole value range below 120 binary orders.
//   ROBUST = true : anchored at the LARGER end value, descending with ratio r <= 1, so an underflow can only
//                   ever drop terms that are negligible against the anchor (anchoring at f_0 would lose
//                   the whole interval when a steep ascending slope flushes e_0); two more instructions.
// The arg-extrema are tracked on the (monotone) scaled values.
template <bool ALL4, bool ROBUST, typename Track>
__device__ __forceinline__ void pass2_x4(const SliceY& sl, float neg_ml, int Dc, int D, float& s, float& ws, Track& trk) {
  float cs0 = sl.template scaled<ALL4>(0, neg_ml);
  {   // d' = 0, 1 sit on slice 0
    const float e = fast_exp2(cs0);
    s = e + e; ws = e;
    trk.update(cs0, 0);
  }
  float base = ROBUST ? 3.5f : 2.f;   // first fine index of the interval, 4k + 2 (ROBUST: + 1.5)
  for (int k = 0; k + 1 < Dc; ++k) {
    const float cs1 = sl.template scaled<ALL4>(k + 1, neg_ml);
    const float dl = cs1 - cs0;
    if constexpr (ROBUST) {
      const float fs = fmaf(0.375f, fabsf(dl), fmaf(0.5f, dl, cs0));        // max(f_0, f_3)
      const float g0 = fast_exp2(fs), r = fast_exp2(-0.25f * fabsf(dl));
      const float g1 = g0 * r, g2 = g1 * r, g3 = g2 * r;
      const float S = (g0 + g1) + (g2 + g3);
      const float Tn = fmaf(3.f, g3, fmaf(2.f, g2, g1));                    // sum_i i * g_i
      // ascending slope (dl >= 0): e_j = g_{3-j}, sum_j (b+j) e_j = (b+3) S - Tn;  descending: b S + Tn
      const float sg = __uint_as_float(0xBF800000u ^ (__float_as_uint(dl) & 0x80000000u));   // -1 : +1
      s += S;
      ws = fmaf(fmaf(-1.5f, sg, base), S, ws);
      ws = fmaf(sg, Tn, ws);
    } else {
      const float e0 = fast_exp2(fmaf(0.125f, dl, cs0)), q = fast_exp2(0.25f * dl);
      const float e1 = e0 * q, e2 = e1 * q, e3 = e2 * q;
      const float S = (e0 + e1) + (e2 + e3);
      const float Tm = fmaf(3.f, e3, fmaf(2.f, e2, e1));                    // sum_j j * e_j
      s += S;
      ws = fmaf(base, S, ws) + Tm;
    }
    // the four fine values of an interval are monotone: only its ends can be extrema -- the far end on a strict
    // slope, the first value otherwise (flat interval: first index; NaN slice: the interval's first value is the
    // first NaN)
    const int d0 = 4 * k + 2;
    const float flo = fmaf(0.125f, dl, cs0), fhi = fmaf(0.875f, dl, cs0);
    trk.update_max(dl > 0.f ? fhi : flo, dl > 0.f ? d0 + 3 : d0);
    trk.update_min(dl < 0.f ? fhi : flo, dl < 0.f ? d0 + 3 : d0);
    base += 4.f;
    cs0 = cs1;
  }
  {   // d' = D-2, D-1 sit on the last slice
    const float e = fast_exp2(cs0);
    s += e + e;
    ws = fmaf((float)(2 * D - 3), e, ws);
    trk.update(cs0, D - 2);
  }
}

// The same pass for TWO horizontally adjacent pixels per thread in packed fp32 (sm_100's FFMA2 / FMUL2 / FADD2: two
// lanes of arithmetic per issue slot -- the one-pixel loop above is bound by instruction issue, 21 instructions per
// interval of which 17 are fp32 arithmetic).  Per interval and pixel pair: two 64-bit LDS, 17 packed operations and four
// ex2, the same operations in the same order as pass2_x4<true, false> on each half.  The 256 threads of a CTA are 128
// pixel pairs x two halves of the disparity range: intervals [kb, ke) here, the caller adds the two halves' sums.
struct PairY {
  const float* p0;   // rows + fy0 * 32 + 2 * (pair column): both pixels share the coarse rows (same y)
  const float* p1;
  float2 w0l, w1l;   // wy * log2(e), the same in both halves
  __device__ __forceinline__ float2 scaled(int k, float2 neg_ml) const {
    const float2 a = *reinterpret_cast<const float2*>(p0 + k * (4 * kTX));
    const float2 b = *reinterpret_cast<const float2*>(p1 + k * (4 * kTX));
    return __ffma2_rn(w0l, a, __ffma2_rn(w1l, b, neg_ml));
  }
};
__device__ __forceinline__ void pass2_x4_pair(const PairY& sl, float2 neg_ml, int kb, int ke, float2& s, float2& ws) {
  const float2 c8 = make_float2(0.125f, 0.125f), c4 = make_float2(0.25f, 0.25f), m1 = make_float2(-1.f, -1.f);
  const float2 two = make_float2(2.f, 2.f), three = make_float2(3.f, 3.f), four = make_float2(4.f, 4.f);
  float2 cs0 = sl.scaled(kb, neg_ml);
  float2 base = make_float2((float)(4 * kb + 2), (float)(4 * kb + 2));   // first fine index of the interval
#pragma unroll 4
  for (int k = kb; k < ke; ++k) {
    const float2 cs1 = sl.scaled(k + 1, neg_ml);
    const float2 dl = __ffma2_rn(cs0, m1, cs1);                           // cs1 - cs0, exactly
    const float2 f0 = __ffma2_rn(c8, dl, cs0), fq = __fmul2_rn(c4, dl);
    float2 e0, q;
    e0.x = fast_exp2(f0.x); e0.y = fast_exp2(f0.y);
    q.x = fast_exp2(fq.x); q.y = fast_exp2(fq.y);
    const float2 e1 = __fmul2_rn(e0, q), e2 = __fmul2_rn(e1, q), e3 = __fmul2_rn(e2, q);
    const float2 S = __fadd2_rn(__fadd2_rn(e0, e1), __fadd2_rn(e2, e3));
    const float2 Tm = __ffma2_rn(three, e3, __ffma2_rn(two, e2, e1));     // sum_j j * e_j
    s = __fadd2_rn(s, S);
    ws = __fadd2_rn(__ffma2_rn(base, S, ws), Tm);
    base = __fadd2_rn(base, four);
    cs0 = cs1;
  }
}

// exact maximum over the FINE values of a pixel (linear inside an interval: attained at an interval end)
template <bool FAST4, bool ALL4>
__device__ __forceinline__ float fine_max(const SliceY& sl, const TailSmem& sm, const TailGeom& g) {
  float c0 = ALL4 ? sl.get4(0) : sl.get(0);
  float M = c0;                                    // d' = 0 sits on slice 0 (and the last ones on slice Dc-1)
  for (int k = 0; k + 1 < g.Dc; ++k) {
    const float c1 = ALL4 ? sl.get4(k + 1) : sl.get(k + 1);
    if constexpr (FAST4) {
      M = fmaxf(M, fmaf(0.375f, fabsf(c1 - c0), 0.5f * (c0 + c1)));
    } else {
      const float dl = c1 - c0;
      const int dend = sm.dstart[k + 1];
      for (int d = sm.dstart[k]; d < dend; ++d) M = fmaxf(M, fmaf(sm.w1tab[d], dl, c0));
    }
    c0 = c1;
  }
  return fmaxf(M, c0);
}

template <typename T, bool FAST4, bool ALL4, bool WANT_ARG>
__global__ void __launch_bounds__(kNT)
upsample_regress_fwd_kernel(const T* __restrict__ cost, T* __restrict__ soft, int64_t* __restrict__ amin,
                            int64_t* __restrict__ amax, float* __restrict__ lse, float* __restrict__ expect, TailGeom g) {
  extern __shared__ __align__(16) float smem_f[];
  const TailSmem sm(smem_f, g);
  const int b = blockIdx.z;
  const int x = blockIdx.x * kTX + (threadIdx.x & (kTX - 1)), y = blockIdx.y * kTY + threadIdx.x / kTX;
  const int cy0 = ALL4 ? lin_index_x4(blockIdx.y * kTY, g.Hc).i0 : lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = ALL4 ? lin_index_x4(blockIdx.x * kTX, g.Wc).i0 : lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  if constexpr (ALL4) stage_tile_all4(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  else stage_tile(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);

  if constexpr (ALL4 && !WANT_ARG) {
    // The x4 head without arg-extrema (what MobileStereoNetV4 evaluates): pixel pairs in packed fp32 (pass2_x4_pair).
    // Thread -> (pair of columns 2 * txp, 2 * txp + 1 of row py, half of the disparity intervals); W % 4 == 0, so a pair
    // is inside the image or outside it as a whole.  A tile with a pixel whose value range could flush (or with NaNs)
    // takes the one-pixel-per-thread path below as a whole.
    const int pi = threadIdx.x & (kNT / 2 - 1), half = threadIdx.x / (kNT / 2);
    const int txp = pi & 15, py = blockIdx.y * kTY + (pi >> 4), px = blockIdx.x * kTX + 2 * txp;
    const bool pvalid = px < g.W && py < g.H;
    const Lin ly = lin_index_x4(py, g.Hc);
    const int o0 = min(ly.i0 - cy0, 3) * kTX + 2 * txp, o1 = min(ly.i1 - cy0, 3) * kTX + 2 * txp;
    const float2 ua = *reinterpret_cast<const float2*>(sm.rowmax + o0), ub = *reinterpret_cast<const float2*>(sm.rowmax + o1);
    const float2 la = *reinterpret_cast<const float2*>(sm.rowmax + 4 * kTX + o0),
                 lb = *reinterpret_cast<const float2*>(sm.rowmax + 4 * kTX + o1);
    const float2 M2 = make_float2(fmaxf(ua.x, ub.x), fmaxf(ua.y, ub.y));
    const bool wide = !((M2.x - fminf(la.x, lb.x)) * kLog2e < 120.f) || !((M2.y - fminf(la.y, lb.y)) * kLog2e < 120.f);
    if (!__syncthreads_or(pvalid && wide)) {
      PairY pr;
      pr.p0 = sm.rows + o0; pr.p1 = sm.rows + o1;
      pr.w0l = make_float2(ly.w0 * kLog2e, ly.w0 * kLog2e); pr.w1l = make_float2(ly.w1 * kLog2e, ly.w1 * kLog2e);
      const float2 nml = make_float2(-M2.x * kLog2e, -M2.y * kLog2e);
      const int kmid = min(g.Dc >> 1, g.Dc - 1);
      float2 s2 = make_float2(0.f, 0.f), ws2 = s2;
      if (pvalid) {
        if (half == 0) {   // d' = 0, 1 sit on slice 0
          const float2 c = pr.scaled(0, nml);
          const float ex = fast_exp2(c.x), ey = fast_exp2(c.y);
          s2 = make_float2(ex + ex, ey + ey); ws2 = make_float2(ex, ey);
          pass2_x4_pair(pr, nml, 0, kmid, s2, ws2);
        } else {           // d' = D-2, D-1 sit on the last slice
          pass2_x4_pair(pr, nml, kmid, g.Dc - 1, s2, ws2);
          const float2 c = pr.scaled(g.Dc - 1, nml);
          const float ex = fast_exp2(c.x), ey = fast_exp2(c.y), wl = (float)(2 * g.D - 3);
          s2.x += ex + ex; s2.y += ey + ey;
          ws2.x = fmaf(wl, ex, ws2.x); ws2.y = fmaf(wl, ey, ws2.y);
        }
      }
      float4* xch = reinterpret_cast<float4*>(sm.part);            // free since the end of the staging
      xch[threadIdx.x] = make_float4(s2.x, s2.y, ws2.x, ws2.y);
      __syncthreads();
      if (pvalid && ((pi >> 5) & 1) == half) {                     // whole warps: each half finishes half of the pairs
        const float4 ot = xch[threadIdx.x ^ (kNT / 2)];
        const float sx = s2.x + ot.x, sy = s2.y + ot.y;
        const float Ex = (ws2.x + ot.z) / sx, Ey = (ws2.y + ot.w) / sy;
        const int64_t o = ((int64_t)b * g.H + py) * g.W + px;
        if (soft) { soft[o] = from_f<T>(Ex); soft[o + 1] = from_f<T>(Ey); }
        if (expect) { expect[o] = Ex; expect[o + 1] = Ey; }        // (scalar stores: the caller's planes need not be 8-byte aligned)
        if (lse) { lse[o] = M2.x + __logf(sx); lse[o + 1] = M2.y + __logf(sy); }
      }
      return;
    }
  }
  if (x >= g.W || y >= g.H) return;

  const SliceY sl(sm, g, y, cy0);
  // pass 1: the softmax stabiliser M.  The general paths take the exact maximum of the pixel's fine values.
  // The all-x4 path takes the bound of stage_tile_all4 (no pass over the slices) whenever the pixel's whole
  // value range is below 120 binary orders -- every ordinary cost volume -- so that nothing can flush and
  // the short progression is safe; wider ranges (or NaNs) take the exact maximum and the robust progression.
  float s, ws;
  typename std::conditional<WANT_ARG, ArgTrack, NoTrack>::type trk;
  float M;
  if constexpr (FAST4) {
    bool robust = true;
    if constexpr (ALL4) {
      M = sl.bound(sm);
      robust = !((M - sl.lower_bound(sm)) * kLog2e < 120.f);
    }
    if (robust) {
      M = fine_max<true, ALL4>(sl, sm, g);
      pass2_x4<ALL4, true>(sl, -M * kLog2e, g.Dc, g.D, s, ws, trk);
    } else {
      pass2_x4<ALL4, false>(sl, -M * kLog2e, g.Dc, g.D, s, ws, trk);
    }
  } else {
    M = fine_max<false, false>(sl, sm, g);
    const float Ml = M * kLog2e;
    s = 0.f; ws = 0.f;
    float cs0 = fmaf(sl.get(0), kLog2e, -Ml);
    for (int k = 0; k < g.Dc; ++k) {
      const float cs1 = (k + 1 < g.Dc) ? fmaf(sl.get(k + 1), kLog2e, -Ml) : cs0;
      const float dl = cs1 - cs0;
      const int dend = sm.dstart[k + 1];
      for (int d = sm.dstart[k]; d < dend; ++d) {
        const float f = fmaf(sm.w1tab[d], dl, cs0);
        const float e = fast_exp2(f);
        s += e;
        ws = fmaf((float)d, e, ws);
        trk.update(f, d);
      }
      cs0 = cs1;
    }
  }
  const int64_t o = ((int64_t)b * g.H + y) * g.W + x;
  const float E = ws / s;
  if (soft) soft[o] = from_f<T>(E);
  if (expect) expect[o] = E;
  if (lse) lse[o] = M + __logf(s);
  if constexpr (WANT_ARG) {
    if (amin) amin[o] = trk.mini;
    if (amax) amax[o] = trk.maxi;
  }
}

// ==================================================================================== backward
// General ratios, stage 1: per fine pixel, the gradient with respect to its slice values c_k -> wsp (B,Dc,H,W) fp32.
//   p(d') = exp(f(d') - lse);  gf = g * p * (d' - E);  gc[i0] += (1-w) gf;  gc[i1] += w gf.
// Deterministic (no atomics): intervals are visited in order, so slice k is complete once
// interval k has been processed.
template <typename T>
__global__ void __launch_bounds__(kNT)
upsample_regress_bwd_cols_kernel(const T* __restrict__ gout, const T* __restrict__ cost,
                                 const float* __restrict__ expect, const float* __restrict__ lse,
                                 float* __restrict__ wsp, TailGeom g) {
  extern __shared__ __align__(16) float smem_f[];
  const TailSmem sm(smem_f, g);
  const int b = blockIdx.z;
  const int x = blockIdx.x * kTX + (threadIdx.x & (kTX - 1)), y = blockIdx.y * kTY + threadIdx.x / kTX;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_tile(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  if (x >= g.W || y >= g.H) return;

  const SliceY sl(sm, g, y, cy0);
  const int64_t o = ((int64_t)b * g.H + y) * g.W + x;
  const float go = to_f(gout[o]), E = expect[o], l2 = lse[o] * kLog2e;
  const int64_t plane = (int64_t)g.H * g.W;
  float* __restrict__ col = wsp + (int64_t)b * g.Dc * plane + (int64_t)y * g.W + x;

  float cs0 = sl.template scaled<false>(0, -l2);
  float acc0 = 0.f;   // gradient of slice k accumulated so far
  if (g.fast4) {
    {
      const float p = fast_exp2(cs0);
      acc0 = go * p * ((0.f - E) + (1.f - E));
    }
    float base = 2.f;
    for (int k = 0; k + 1 < g.Dc; ++k) {
      const float cs1 = sl.template scaled<false>(k + 1, -l2);
      const float dl = cs1 - cs0;
      float acc1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = 0.125f + 0.25f * j;
        const float p = fast_exp2(fmaf(w, dl, cs0));
        const float gf = go * p * ((base + (float)j) - E);
        acc0 = fmaf(1.f - w, gf, acc0);
        acc1 = fmaf(w, gf, acc1);
      }
      __stcs(col + (int64_t)k * plane, acc0);
      acc0 = acc1;
      base += 4.f;
      cs0 = cs1;
    }
    {
      const float p = fast_exp2(cs0);
      acc0 += go * p * (((float)(g.D - 2) - E) + ((float)(g.D - 1) - E));
      __stcs(col + (int64_t)(g.Dc - 1) * plane, acc0);
    }
  } else {
    for (int k = 0; k < g.Dc; ++k) {
      const float cs1 = (k + 1 < g.Dc) ? fmaf(sl.get(k + 1), kLog2e, -l2) : cs0;
      const float dl = cs1 - cs0;
      float acc1 = 0.f;
      const int dend = sm.dstart[k + 1];
      for (int d = sm.dstart[k]; d < dend; ++d) {
        const float w = sm.w1tab[d];
        const float p = fast_exp2(fmaf(w, dl, cs0));
        const float gf = go * p * ((float)d - E);
        acc0 = fmaf(1.f - w, gf, acc0);
        acc1 = fmaf(w, gf, acc1);      // for k == Dc-1, i1 == i0: dl == 0 and acc1 is folded below
      }
      if (k + 1 == g.Dc) acc0 += acc1;
      col[(int64_t)k * plane] = acc0;
      acc0 = acc1;
      cs0 = cs1;
    }
  }
}

// ---- x4 x4 x4 head: stage 1 with the transposed bilinear interpolation done inside the CTA.
// The general stage 1 above writes one fp32 value per (slice, FINE pixel) -- a (B,Dc,H,W) workspace, 16x the
// gradient itself -- and stage 2 gathers 64 taps per coarse voxel from it: at (8,48,96,312) -> (8,192,384,1248) that is
// 1.47 GB of HBM traffic and a 358 us gather next to a 367 us stage 1.  Here a CTA reduces its 32 x 8 fine pixels
// to the 4 x 10 coarse voxels they touch, eight slices at a time, with the x4 rule's constant weights: fine
// columns 4j .. 4j+3 give (3 v0 + v1)/8 to coarse column j-1, (5 v0 + 7 v1 + 7 v2 + 5 v3)/8 to j and (v2 + 3 v3)/8 to
// j+1, the same along y.  The footprint is addressed UNCLAMPED (origin (8 bx - 1, 2 by - 1), possibly -1): the
// align_corners=False clamp at the image border moves a weight from voxel -1 to voxel 0 (or from Wc to Wc-1), which
// the second kernel does by folding those cells -- so the tile kernel has no border cases at all, and out-of-image
// pixels just contribute zeros.  The CTA writes 40 partial sums per slice; the second kernel adds, for every coarse
// voxel, the cells of the <= 4 tiles that hold it, in a fixed order.  Deterministic, no atomics; the workspace is
// (B, tiles, Dc, 40): 6.4x smaller than (B,Dc,H,W).
// The exponentials of an interval's four fine values come from two ex2 and three multiplies (geometric
// progression, as in the forward pass) whenever the pixel's value range leaves 2^-126 out of reach.
constexpr int kKB = 8;            // slices per reduction batch
constexpr int kAP = kTX + 4;      // pitch of a staged gradient row (16-byte aligned rows)
constexpr int kTileCells = 40;    // 4 x 10 coarse voxels per tile and slice
constexpr int kBwdExtra = kKB * kTY * kAP + kKB * kTY * 10;   // A | B (floats)

// one interval: slices k, k + 1 = cs0, cs1; adds the four fine values' gradient to acc0 (slice k), returns slice k+1's
template <bool ROBUST>
__device__ __forceinline__ float tail_bwd_interval(float cs0, float cs1, float u, float go, float& acc0) {
  const float dl = cs1 - cs0;
  float p0, p1, p2, p3;
  p0 = fast_exp2(fmaf(0.125f, dl, cs0));
  if constexpr (!ROBUST) {
    const float q = fast_exp2(0.25f * dl);
    p1 = p0 * q; p2 = p1 * q; p3 = p2 * q;
  } else {
    p1 = fast_exp2(fmaf(0.375f, dl, cs0)); p2 = fast_exp2(fmaf(0.625f, dl, cs0)); p3 = fast_exp2(fmaf(0.875f, dl, cs0));
  }
  const float g0 = p0 * u, g1 = p1 * (u + go), g2 = p2 * fmaf(2.f, go, u), g3 = p3 * fmaf(3.f, go, u);
  acc0 = fmaf(0.875f, g0, acc0); acc0 = fmaf(0.625f, g1, acc0); acc0 = fmaf(0.375f, g2, acc0); acc0 = fmaf(0.125f, g3, acc0);
  return fmaf(0.875f, g3, fmaf(0.625f, g2, fmaf(0.375f, g1, 0.125f * g0)));
}

// slices [k0, k0 + nk) of one pixel -> column (ty, tx) of the staged gradients.  MODE 0: eight whole intervals;
// MODE 1: seven intervals and the last slice (the final batch when Dc % 8 == 0); MODE 2: any nk, tested per slice
template <bool ROBUST, int MODE>
__device__ __forceinline__ void tail_bwd_batch(const SliceY& sl, float* __restrict__ a, int k0, int nk, int Dc, int D,
                                               float l2, float E, float go, bool valid, float& cs0, float& acc0, float& u) {
#pragma unroll
  for (int kk = 0; kk < kKB; ++kk) {
    if (MODE == 2 && kk >= nk) break;
    float acc1 = 0.f;
    if (MODE == 0 || (MODE == 1 && kk < kKB - 1) || (MODE == 2 && k0 + kk + 1 < Dc)) {
      const float cs1 = sl.template scaled<true>(k0 + kk + 1, -l2);
      acc1 = tail_bwd_interval<ROBUST>(cs0, cs1, u, go, acc0);
      u = fmaf(4.f, go, u);
      cs0 = cs1;
    } else {
      acc0 += go * fast_exp2(cs0) * (((float)(D - 2) - E) + ((float)(D - 1) - E));   // the last two fine values
    }
    a[kk * kTY * kAP] = valid ? acc0 : 0.f;
    acc0 = acc1;
  }
}

template <typename T>
__global__ void __launch_bounds__(kNT)
upsample_regress_bwd_tile_kernel(const T* __restrict__ gout, const T* __restrict__ cost,
                                 const float* __restrict__ expect, const float* __restrict__ lse,
                                 float* __restrict__ part, TailGeom g) {
  extern __shared__ __align__(16) float smem_f[];
  const TailSmem sm(smem_f, g);
  float* sA = sm.rowmax + 2 * 4 * kTX;         // [kKB][kTY][kAP]
  float* sB = sA + kKB * kTY * kAP;            // [kKB][kTY][10]
  const int b = blockIdx.z;
  const int tx = threadIdx.x & (kTX - 1), ty = threadIdx.x / kTX;
  const int x = blockIdx.x * kTX + tx, y = blockIdx.y * kTY + ty;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_tile_all4(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);

  const bool valid = x < g.W && y < g.H;
  const SliceY sl(sm, g, min(y, g.H - 1), cy0);
  const int64_t o = ((int64_t)b * g.H + min(y, g.H - 1)) * g.W + min(x, g.W - 1);
  const float go = valid ? to_f(gout[o]) : 0.f, E = expect[o], l2 = lse[o] * kLog2e;
  // progression only when no fine value of the pixel can come within 2^-126 of flushing: range + log2(D) < 118
  const bool robust = !((sl.bound(sm) - sl.lower_bound(sm)) * kLog2e + __log2f((float)g.D) < 118.f);
  float* __restrict__ ptile = part + ((((int64_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * g.Dc) * kTileCells;
  float* __restrict__ acol = sA + ty * kAP + tx;

  float cs0 = sl.template scaled<true>(0, -l2);
  float acc0 = go * fast_exp2(cs0) * ((0.f - E) + (1.f - E));   // d' = 0, 1 sit on slice 0
  float u = go * (2.f - E);                                      // go * (first fine index of the interval - E)
#pragma unroll 1
  for (int k0 = 0; k0 < g.Dc; k0 += kKB) {
    const int nk = min(kKB, g.Dc - k0);
    if (k0 + kKB < g.Dc) {
      if (!robust) tail_bwd_batch<false, 0>(sl, acol, k0, nk, g.Dc, g.D, l2, E, go, valid, cs0, acc0, u);
      else tail_bwd_batch<true, 0>(sl, acol, k0, nk, g.Dc, g.D, l2, E, go, valid, cs0, acc0, u);
    } else if (nk == kKB) {
      if (!robust) tail_bwd_batch<false, 1>(sl, acol, k0, nk, g.Dc, g.D, l2, E, go, valid, cs0, acc0, u);
      else tail_bwd_batch<true, 1>(sl, acol, k0, nk, g.Dc, g.D, l2, E, go, valid, cs0, acc0, u);
    } else {
      tail_bwd_batch<true, 2>(sl, acol, k0, nk, g.Dc, g.D, l2, E, go, valid, cs0, acc0, u);
    }
    __syncthreads();
    // ---- along x: item (row = kk * 8 + ty, q) reads its fine quad and exchanges with its neighbours in the warp
#pragma unroll 1
    for (int it = threadIdx.x; it < nk * kTY * 8; it += kNT) {
      const int q = it & 7, row = it >> 3;
      const float4 v = *reinterpret_cast<const float4*>(sA + row * kAP + 4 * q);
      const float sm_ = fmaf(0.375f, v.x, 0.125f * v.y);
      const float sc = fmaf(0.625f, v.x + v.w, 0.875f * (v.y + v.z));
      const float sp = fmaf(0.375f, v.w, 0.125f * v.z);
      // footprint column fx (coarse 8 bx - 1 + fx) = sm[fx] + sc[fx - 1] + sp[fx - 2]
      const float sc1 = __shfl_up_sync(0xffffffffu, sc, 1, 8), sp1 = __shfl_up_sync(0xffffffffu, sp, 1, 8);
      const float sp2 = __shfl_up_sync(0xffffffffu, sp, 2, 8);
      float* brow = sB + row * 10;
      brow[q] = sm_ + (q >= 1 ? sc1 : 0.f) + (q >= 2 ? sp2 : 0.f);
      if (q == 7) { brow[8] = sc + sp1; brow[9] = sp; }
    }
    __syncthreads();
    // ---- along y: item (kk, fx) reads the column's 8 rows; footprint row fy (coarse 2 by - 1 + fy)
#pragma unroll 1
    for (int it = threadIdx.x; it < nk * 10; it += kNT) {
      const int kk = it / 10, fx = it - kk * 10;
      const float* bb = sB + kk * kTY * 10 + fx;
      const float r0 = bb[0], r1 = bb[10], r2 = bb[20], r3 = bb[30], r4 = bb[40], r5 = bb[50], r6 = bb[60], r7 = bb[70];
      const float m0 = fmaf(0.375f, r0, 0.125f * r1), c0 = fmaf(0.625f, r0 + r3, 0.875f * (r1 + r2)), p0 = fmaf(0.375f, r3, 0.125f * r2);
      const float m1 = fmaf(0.375f, r4, 0.125f * r5), c1 = fmaf(0.625f, r4 + r7, 0.875f * (r5 + r6)), p1 = fmaf(0.375f, r7, 0.125f * r6);
      float* o4 = ptile + (int64_t)(k0 + kk) * kTileCells + fx;
      o4[0] = m0; o4[10] = c0 + m1; o4[20] = p0 + c1; o4[30] = p1;
    }
    // (the next batch's A stores do not touch B, and its first barrier orders its B stores after these reads)
  }
}

// stage 2 of the tile form: every coarse voxel adds the cells that hold it -- tile (bx, by) holds coarse columns
// 8 bx - 1 .. 8 bx + 8 and rows 2 by - 1 .. 2 by + 2 -- plus, at the image border, the cells one step outside
// (the clamp of align_corners=False).  A thread finds the <= 4 (border: <= 16) cells of its (y, x) once and then walks
// its share of the slices (the cell offsets do not depend on k).  grid = (ceil(Wc / 64), Hc, B * ksplit): the split of
// the slices only adds loads in flight (the kernel is bound by their latency, not by issue or bandwidth).
template <typename T>
__global__ void __launch_bounds__(64)
upsample_regress_bwd_combine_kernel(const float* __restrict__ part, T* __restrict__ gcost, TailGeom g, int tiles_x,
                                    int tiles_y, int ksplit) {
  const int xc = blockIdx.x * 64 + threadIdx.x, yc = blockIdx.y, b = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  if (xc >= g.Wc) return;
  const int kbeg = g.Dc * ks / ksplit, kend = g.Dc * (ks + 1) / ksplit;
  const float* __restrict__ pb = part + (int64_t)b * tiles_y * tiles_x * g.Dc * kTileCells;
  T* __restrict__ out = gcost + ((int64_t)b * g.Dc * g.Hc + yc) * g.Wc + xc;
  const int64_t plane = (int64_t)g.Hc * g.Wc;
  const int tstride = g.Dc * kTileCells;
  // cell offsets (tile * Dc * 40 + fy * 10 + fx) in the fixed order (row, column) ascending; -1 = none
  auto cells_of = [&](int yy, int xx, int (&off)[4]) {
    const int fy0 = (yy + 1) & 1, fx0 = (xx + 1) & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int fy = fy0 + 2 * (j >> 1), fx = fx0 + 8 * (j & 1);
      const int by = (yy + 1 - fy) >> 1, bx = (xx + 1 - fx) >> 3;
      off[j] = (fx > 9 || by < 0 || by >= tiles_y || bx < 0 || bx >= tiles_x) ? -1 : (by * tiles_x + bx) * tstride + fy * 10 + fx;
    }
  };
  const bool border = yc == 0 || yc == g.Hc - 1 || xc == 0 || xc == g.Wc - 1;
  if (!border) {
    int off[4];
    cells_of(yc, xc, off);
#pragma unroll 4
    for (int k = kbeg; k < kend; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (off[j] >= 0) acc += __ldg(pb + off[j] + k * kTileCells);
      out[k * plane] = from_f<T>(acc);
    }
    return;
  }
  // border voxels: the unclamped coordinates folded onto this voxel: itself, -1 onto 0, Hc onto Hc - 1 (same along x)
  const int ny = 1 + (yc == 0) + (yc == g.Hc - 1), nx = 1 + (xc == 0) + (xc == g.Wc - 1);
  for (int k = kbeg; k < kend; ++k) {
    float acc = 0.f;
    for (int iy = 0; iy < ny; ++iy) {
      const int yy = iy == 0 ? yc : (iy == 1 && yc == 0 ? -1 : g.Hc);
      for (int ix = 0; ix < nx; ++ix) {
        const int xx = ix == 0 ? xc : (ix == 1 && xc == 0 ? -1 : g.Wc);
        int off[4];
        cells_of(yy, xx, off);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (off[j] >= 0) acc += __ldg(pb + off[j] + k * kTileCells);
      }
    }
    out[k * plane] = from_f<T>(acc);
  }
}

// range of fine indices whose (i0 or i1) can equal coarse index ic (conservative; exact test inside)
__device__ __forceinline__ void fine_range(int ic, float scale, int n_out, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = max(0, (int)floorf(((float)ic - 0.5f) * inv - 0.5f) - 1);
  hi = min(n_out - 1, (int)ceilf(((float)ic + 1.5f) * inv - 0.5f) + 1);
}

// stage 2: transposed bilinear gather, one thread per coarse element
template <typename T>
__global__ void __launch_bounds__(256)
upsample_regress_bwd_gather_kernel(const float* __restrict__ wsp, T* __restrict__ gcost, int64_t total,
                                   TailGeom g) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int xc = (int)(i % g.Wc);
  const int yc = (int)((i / g.Wc) % g.Hc);
  const int64_t bk = i / ((int64_t)g.Wc * g.Hc);
  int ylo, yhi, xlo, xhi;
  fine_range(yc, g.sh, g.H, ylo, yhi);
  fine_range(xc, g.sw, g.W, xlo, xhi);
  const float* __restrict__ src = wsp + bk * (int64_t)g.H * g.W;
  float acc = 0.f;
  if (g.fastx && g.fasty && yc >= 1 && yc <= g.Hc - 2 && xc >= 1 && xc <= g.Wc - 2) {
    // x4 in both axes, interior voxel: the 8 x 8 fine pixels 4c-2 .. 4c+5 with the constant separable
    // weights {1,3,5,7,7,5,3,1}/8 (the transposed lin_index table) -- no index arithmetic per tap
    const float* p = src + (int64_t)(4 * yc - 2) * g.W + (4 * xc - 2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float2* r = reinterpret_cast<const float2*>(p + (int64_t)t * g.W);
      const float2 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2), d = __ldg(r + 3);
      const float racc = 0.125f * (a.x + d.y) + 0.375f * (a.y + d.x) + 0.625f * (b.x + c.y) + 0.875f * (b.y + c.x);
      const float wy = t < 4 ? 0.125f + 0.25f * t : 0.875f - 0.25f * (t - 4);
      acc = fmaf(wy, racc, acc);
    }
    gcost[i] = from_f<T>(acc);
    return;
  }
  for (int y = ylo; y <= yhi; ++y) {
    const Lin ly = lin_index(y, g.sh, g.Hc);
    if (ly.i0 != yc && ly.i1 != yc) continue;
    const float wy = (ly.i0 == yc ? ly.w0 : 0.f) + (ly.i1 == yc ? ly.w1 : 0.f);
    float racc = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      const Lin lx = lin_index(x, g.sw, g.Wc);
      const float wx = (lx.i0 == xc ? lx.w0 : 0.f) + (lx.i1 == xc ? lx.w1 : 0.f);
      if (lx.i0 == xc || lx.i1 == xc) racc = fmaf(wx, __ldg(src + (int64_t)y * g.W + x), racc);
    }
    acc = fmaf(wy, racc, acc);
  }
  gcost[i] = from_f<T>(acc);
}

static bool grid_ok(int64_t blocks) { return blocks >= 0 && blocks <= 2147483647LL; }

static int make_geom(int64_t Dc, int64_t Hc, int64_t Wc, int64_t D, int64_t H, int64_t W, TailGeom& g,
                     size_t& smem) {
  if (Dc <= 0 || Hc <= 0 || Wc <= 0 || D <= 0 || H <= 0 || W <= 0) return RSM_ERR_INVALID_SHAPE;
  if (Dc > (1 << 20) || Hc > (1 << 20) || Wc > (1 << 20) || D > (1 << 20) || H > (1 << 20) || W > (1 << 20))
    return RSM_ERR_INVALID_SHAPE;
  g.Dc = (int)Dc; g.Hc = (int)Hc; g.Wc = (int)Wc; g.D = (int)D; g.H = (int)H; g.W = (int)W;
  g.sd = (float)Dc / (float)D; g.sh = (float)Hc / (float)H; g.sw = (float)Wc / (float)W;
  g.FH = (int)fminf((float)Hc, ceilf(kTY * g.sh) + 2.f);
  g.FW = (int)fminf((float)Wc, ceilf(kTX * g.sw) + 2.f);
  g.fast4 = (D == 4 * Dc) ? 1 : 0;
  g.fastx = (W == 4 * Wc) ? 1 : 0;
  g.fasty = (H == 4 * Hc) ? 1 : 0;
  g.all4 = g.fast4 && g.fastx && g.fasty && g.FH == 4 && g.FW == 10 && Dc * Hc * Wc < 2147483647LL;
  smem = tail_smem_bytes(g);
  if (smem > 200 * 1024) return RSM_ERR_UNSUPPORTED_CONFIG;
  return RSM_OK;
}

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

template <typename T, bool FAST4, bool ALL4, bool WANT_ARG>
static int launch_tail_fwd(const void* cost, const rsm_regress_out& out, const TailGeom& g, size_t smem, dim3 grid,
                           cudaStream_t st) {
  auto k = upsample_regress_fwd_kernel<T, FAST4, ALL4, WANT_ARG>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<grid, kNT, smem, st>>>((const T*)cost, (T*)out.soft, out.argmin, out.argmax, out.lse, out.expect, g);
  return finish_launch("rsm_upsample_regress_fwd");
}

extern "C" int rsm_upsample_regress_fwd(const void* cost, int64_t B, int64_t Dc, int64_t Hc, int64_t Wc,
                                        int64_t D, int64_t H, int64_t W, int dtype, rsm_regress_out out,
                                        int device, void* stream) {
  if (B < 0) return RSM_ERR_INVALID_SHAPE;
  TailGeom g;
  size_t smem;
  if (int rc = make_geom(Dc, Hc, Wc, D, H, W, g, smem)) return rc;
  if (B == 0) return RSM_OK;
  if (!cost) return RSM_ERR_NULL_POINTER;
  if (B > 65535) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  const dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)B);
  if (grid.y > 65535) return RSM_ERR_INVALID_SHAPE;
  const bool want_arg = out.argmin || out.argmax;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    if (g.all4) return want_arg ? launch_tail_fwd<T, true, true, true>(cost, out, g, smem, grid, st)
                                : launch_tail_fwd<T, true, true, false>(cost, out, g, smem, grid, st);
    if (g.fast4) return want_arg ? launch_tail_fwd<T, true, false, true>(cost, out, g, smem, grid, st)
                                 : launch_tail_fwd<T, true, false, false>(cost, out, g, smem, grid, st);
    return want_arg ? launch_tail_fwd<T, false, false, true>(cost, out, g, smem, grid, st)
                    : launch_tail_fwd<T, false, false, false>(cost, out, g, smem, grid, st);
  });
}

extern "C" int64_t rsm_upsample_regress_bwd_workspace(int64_t B, int64_t Dc, int64_t H, int64_t W) {
  if (B < 0 || Dc < 0 || H < 0 || W < 0) return -1;
  return B * Dc * H * W * (int64_t)sizeof(float);
}

extern "C" int rsm_upsample_regress_bwd(const void* gout, const void* cost, const float* expect,
                                        const float* lse, void* gcost, void* workspace, int64_t B,
                                        int64_t Dc, int64_t Hc, int64_t Wc, int64_t D, int64_t H,
                                        int64_t W, int dtype, int device, void* stream) {
  if (B < 0) return RSM_ERR_INVALID_SHAPE;
  TailGeom g;
  size_t smem;
  if (int rc = make_geom(Dc, Hc, Wc, D, H, W, g, smem)) return rc;
  if (B == 0) return RSM_OK;
  if (!gout || !cost || !expect || !lse || !gcost || !workspace) return RSM_ERR_NULL_POINTER;
  if (B > 65535) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    if (g.all4) {
      // x4 x4 x4 head: per-tile partial sums (B, tiles, Dc, 40) + combine
      const size_t smem_t = smem + kBwdExtra * sizeof(float);
      const dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)B);
      if (grid.y > 65535 || Hc > 65535 || smem_t > 200 * 1024 || (int64_t)grid.x * grid.y * Dc * kTileCells > 2147483647LL)
        return (int)RSM_ERR_INVALID_SHAPE;
      auto kt = upsample_regress_bwd_tile_kernel<T>;
      if (smem_t > 48 * 1024) cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
      kt<<<grid, kNT, smem_t, st>>>((const T*)gout, (const T*)cost, expect, lse, (float*)workspace, g);
      if (int rc = finish_launch("rsm_upsample_regress_bwd(tiles)")) return rc;
      // enough CTAs for ~4 waves of 32 x 64 threads per SM
      int64_t ksplit = ceil_div((int64_t)4 * kNumSMs * 32, ceil_div(Wc, 64) * Hc * B);
      ksplit = ksplit < 1 ? 1 : (ksplit > Dc ? Dc : ksplit);
      if (B * ksplit > 65535) ksplit = 1;
      if (B > 65535) return (int)RSM_ERR_INVALID_SHAPE;
      upsample_regress_bwd_combine_kernel<T><<<dim3((unsigned)ceil_div(Wc, 64), (unsigned)Hc, (unsigned)(B * ksplit)), 64, 0, st>>>(
          (const float*)workspace, (T*)gcost, g, (int)grid.x, (int)grid.y, (int)ksplit);
      return finish_launch("rsm_upsample_regress_bwd(combine)");
    }
    auto k = upsample_regress_bwd_cols_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)B);
    if (grid.y > 65535) return (int)RSM_ERR_INVALID_SHAPE;
    k<<<grid, kNT, smem, st>>>((const T*)gout, (const T*)cost, expect, lse, (float*)workspace, g);
    if (int rc = finish_launch("rsm_upsample_regress_bwd(cols)")) return rc;
    const int64_t total = B * Dc * Hc * Wc;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    upsample_regress_bwd_gather_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const float*)workspace, (T*)gcost, total, g);
    return finish_launch("rsm_upsample_regress_bwd(gather)");
  });
}
