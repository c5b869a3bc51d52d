// MobileStereoNetV4 head, fused: F.interpolate(cost[:,None], [D,H,W], 'trilinear') -> softmax over D
// -> expectation (model/mobile_stereo_net_v4.py:511-518), forward and adjoint, without ever
// materialising the (B,D,H,W) tensor.  The kernel is bound by instruction issue and MUFU (one
// exp per fine disparity), not by HBM, so the work per fine disparity is stripped
// to: one FMA (lerp along d of two pre-scaled coarse slices), one ex2, and the two accumulations.
//
// Per output pixel the 4 bilinear taps of every coarse slice are combined ONCE per slice
// ("slice value" c_k); fine disparities are then visited interval by interval: all d' whose
// source index i0(d') equals k interpolate between c_k and c_{k+1}.  For the x4 head
// (D == 4*Dc) the intervals are regular -- weights 1/8,3/8,5/8,7/8 -- and the inner loop is fully
// unrolled with constant weights; other ratios use per-CTA tables in shared memory.
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

constexpr int kTX = 32, kTY = 8;  // fine-pixel tile of one CTA (256 threads, one pixel each)

struct TailGeom {
  int Dc, Hc, Wc, D, H, W;
  int FH, FW;        // coarse rows / cols a tile can touch (upper bound)
  int fast4;         // D == 4 * Dc
  int cached;        // per-pixel slice values are parked in shared memory (Dc * 256 floats fit)
  float sd, sh, sw;  // in/out scale per axis
};

// shared memory: [ footprint Dc*FH*FW | slices Dc*256 (if cached) | w1tab D | dstart Dc+1 | i0tab D ]
// (tables: generic ratio only)
struct TailSmem {
  float* foot;
  float* slices;   // slices[k * 256 + tid]: bilinear slice value c_k of this thread's pixel
  float* w1tab;
  int* dstart;
  int* i0tab;
  __device__ __forceinline__ TailSmem(float* base, const TailGeom& g) {
    foot = base;
    slices = base + g.Dc * g.FH * g.FW;
    w1tab = slices + (g.cached ? g.Dc * kTX * kTY : 0);
    dstart = reinterpret_cast<int*>(w1tab + g.D);
    i0tab = dstart + g.Dc + 1;
  }
};
static size_t tail_smem_bytes(const TailGeom& g) {
  size_t n = (size_t)g.Dc * g.FH * g.FW;
  if (g.cached) n += (size_t)g.Dc * kTX * kTY;
  if (!g.fast4) n += (size_t)g.D + g.Dc + 1 + g.D;
  return n * sizeof(float);
}

// stage the coarse footprint of this tile as fp32: foot[k][fy][fx], rows cy0.., cols cx0..; and,
// for generic ratios, the per-fine-disparity tables.  Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void stage_tile(const T* __restrict__ cost_b, const TailSmem& sm, const TailGeom& g,
                                           int cy0, int cx0) {
  const int per = g.FH * g.FW;
  const int tot = g.Dc * per;
  const int64_t plane = (int64_t)g.Hc * g.Wc;
  constexpr int U = 8, NT = kTX * kTY;   // U independent loads in flight per thread
  for (int e0 = threadIdx.x; e0 < tot; e0 += U * NT) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * NT;
      v[u] = 0.f;
      if (e < tot) {
        const int k = e / per, r = e - k * per;
        const int fy = r / g.FW, fx = r - fy * g.FW;
        const int cy = min(cy0 + fy, g.Hc - 1), cx = min(cx0 + fx, g.Wc - 1);
        v[u] = to_f(__ldg(cost_b + k * plane + (int64_t)cy * g.Wc + cx));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e0 + u * NT < tot) sm.foot[e0 + u * NT] = v[u];
  }
  if (!g.fast4) {
    for (int d = threadIdx.x; d < g.D; d += kTX * kTY) {
      const Lin ld = lin_index(d, g.sd, g.Dc);
      sm.w1tab[d] = ld.w1;
      sm.i0tab[d] = ld.i0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= g.Dc; k += kTX * kTY) {   // dstart[k] = #{d : i0(d) < k}
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

// bilinear taps of one output pixel inside the staged footprint
struct Taps {
  int o00, o01, o10, o11;
  float wx0, wx1, wy0, wy1;
  __device__ __forceinline__ Taps(int x, int y, int cx0, int cy0, const TailGeom& g) {
    const Lin ly = lin_index(y, g.sh, g.Hc), lx = lin_index(x, g.sw, g.Wc);
    o00 = (ly.i0 - cy0) * g.FW + (lx.i0 - cx0); o01 = (ly.i0 - cy0) * g.FW + (lx.i1 - cx0);
    o10 = (ly.i1 - cy0) * g.FW + (lx.i0 - cx0); o11 = (ly.i1 - cy0) * g.FW + (lx.i1 - cx0);
    wx0 = lx.w0; wx1 = lx.w1; wy0 = ly.w0; wy1 = ly.w1;
  }
  __device__ __forceinline__ float slice(const float* s) const {
    return wy0 * (wx0 * s[o00] + wx1 * s[o01]) + wy1 * (wx0 * s[o10] + wx1 * s[o11]);
  }
};

// slice values of one pixel: evaluated once (pass 1) and, when they fit, parked in shared memory so
// that pass 2 is one LDS per slice instead of 4 LDS + 6 flops
template <bool CACHED>
struct SliceSrc {
  const Taps& tp;
  const float* foot;
  float* mine;   // &slices[tid]
  int per;
  __device__ __forceinline__ float max_and_park(int Dc) const {
    float M = -INFINITY;
    const float* s = foot;
    float* dst = mine;
    for (int k = 0; k < Dc; ++k, s += per, dst += kTX * kTY) {
      const float c = tp.slice(s);
      M = fmaxf(M, c);
      if (CACHED) *dst = c;
    }
    return M;
  }
  __device__ __forceinline__ float get(int k) const {
    return CACHED ? mine[k * (kTX * kTY)] : tp.slice(foot + k * per);
  }
};

struct NoTrack {
  __device__ __forceinline__ void update(float, int) {}
};

// ===================================================================================== forward
template <typename T, bool FAST4, bool WANT_ARG, bool CACHED>
__global__ void __launch_bounds__(kTX * kTY)
upsample_regress_fwd_kernel(const T* __restrict__ cost, T* __restrict__ soft, int64_t* __restrict__ amin,
                            int64_t* __restrict__ amax, float* __restrict__ lse, TailGeom g) {
  extern __shared__ __align__(16) float smem_f[];
  const TailSmem sm(smem_f, g);
  const int b = blockIdx.z;
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x = blockIdx.x * kTX + tx, y = blockIdx.y * kTY + ty;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_tile(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  if (x >= g.W || y >= g.H) return;

  const Taps tp(x, y, cx0, cy0, g);
  const SliceSrc<CACHED> src{tp, sm.foot, sm.slices + threadIdx.x, g.FH * g.FW};
  // pass 1: stabiliser.  Every fine value is a convex combination of slice values, so their max
  // bounds it (and is attained within |c_{k+1}-c_k|/8 for the x4 head).
  const float M = src.max_and_park(g.Dc);
  const float Ml = M * kLog2e;

  // pass 2: intervals in ascending d.  Slices are pre-scaled: cs = c*log2(e) - M*log2(e), so
  // exp(f - M) = ex2(lerp(cs0, cs1)).  The arg-extrema are tracked on the (monotone) scaled values.
  float s = 0.f, ws = 0.f;
  typename std::conditional<WANT_ARG, ArgTrack, NoTrack>::type trk;
  float cs0 = fmaf(src.get(0), kLog2e, -Ml);
  if constexpr (FAST4) {
    {   // d' = 0, 1 sit on slice 0
      const float e = fast_exp2(cs0);
      s = e + e; ws = e;
      trk.update(cs0, 0);
    }
    float base = 2.f;   // first fine index of the interval, 4k + 2
    for (int k = 0; k + 1 < g.Dc; ++k) {
      const float cs1 = fmaf(src.get(k + 1), kLog2e, -Ml);
      const float dl = cs1 - cs0;
      const float f0 = fmaf(0.125f, dl, cs0), f1 = fmaf(0.375f, dl, cs0);
      const float f2 = fmaf(0.625f, dl, cs0), f3 = fmaf(0.875f, dl, cs0);
      const float e0 = fast_exp2(f0), e1 = fast_exp2(f1), e2 = fast_exp2(f2), e3 = fast_exp2(f3);
      const float S = (e0 + e1) + (e2 + e3);
      const float Tm = fmaf(3.f, e3, fmaf(2.f, e2, e1));   // sum_j j * e_j
      s += S;
      ws = fmaf(base, S, ws) + Tm;
      if constexpr (WANT_ARG) {
        const int d0 = 4 * k + 2;
        trk.update(f0, d0); trk.update(f1, d0 + 1); trk.update(f2, d0 + 2); trk.update(f3, d0 + 3);
      }
      base += 4.f;
      cs0 = cs1;
    }
    {   // d' = D-2, D-1 sit on the last slice
      const float e = fast_exp2(cs0);
      s += e + e;
      ws = fmaf((float)(2 * g.D - 3), e, ws);
      trk.update(cs0, g.D - 2);
    }
  } else {
    for (int k = 0; k < g.Dc; ++k) {
      const float cs1 = (k + 1 < g.Dc) ? fmaf(src.get(k + 1), kLog2e, -Ml) : cs0;
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
  if (soft) soft[o] = from_f<T>(ws / s);
  if (lse) lse[o] = M + __logf(s);
  if constexpr (WANT_ARG) {
    if (amin) amin[o] = trk.mini;
    if (amax) amax[o] = trk.maxi;
  }
}

// ==================================================================================== backward
// stage 1: per fine pixel, the gradient with respect to its slice values c_k -> wsp (B,Dc,H,W) fp32.
//   p(d') = exp(f(d') - lse);  gf = g * p * (d' - E);  gc[i0] += (1-w) gf;  gc[i1] += w gf.
// Deterministic (no atomics): intervals are visited in order, so slice k is complete once
// interval k has been processed.
template <typename T>
__global__ void __launch_bounds__(kTX * kTY)
upsample_regress_bwd_cols_kernel(const T* __restrict__ gout, const T* __restrict__ cost,
                                 const T* __restrict__ soft, const float* __restrict__ lse,
                                 float* __restrict__ wsp, TailGeom g) {
  extern __shared__ __align__(16) float smem_f[];
  const TailSmem sm(smem_f, g);
  const int b = blockIdx.z;
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x = blockIdx.x * kTX + tx, y = blockIdx.y * kTY + ty;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_tile(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  if (x >= g.W || y >= g.H) return;

  const Taps tp(x, y, cx0, cy0, g);
  const int per = g.FH * g.FW;
  const int64_t o = ((int64_t)b * g.H + y) * g.W + x;
  const float go = to_f(gout[o]), E = to_f(soft[o]), l2 = lse[o] * kLog2e;
  const int64_t plane = (int64_t)g.H * g.W;
  float* __restrict__ col = wsp + (int64_t)b * g.Dc * plane + (int64_t)y * g.W + x;

  float cs0 = fmaf(tp.slice(sm.foot), kLog2e, -l2);
  float acc0 = 0.f;   // gradient of slice k accumulated so far
  if (g.fast4) {
    {
      const float p = fast_exp2(cs0);
      acc0 = go * p * ((0.f - E) + (1.f - E));
    }
    float base = 2.f;
    for (int k = 0; k + 1 < g.Dc; ++k) {
      const float cs1 = fmaf(tp.slice(sm.foot + (k + 1) * per), kLog2e, -l2);
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
      col[(int64_t)k * plane] = acc0;
      acc0 = acc1;
      base += 4.f;
      cs0 = cs1;
    }
    {
      const float p = fast_exp2(cs0);
      acc0 += go * p * (((float)(g.D - 2) - E) + ((float)(g.D - 1) - E));
      col[(int64_t)(g.Dc - 1) * plane] = acc0;
    }
  } else {
    for (int k = 0; k < g.Dc; ++k) {
      const float cs1 = (k + 1 < g.Dc) ? fmaf(tp.slice(sm.foot + (k + 1) * per), kLog2e, -l2) : cs0;
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
  g.cached = (Dc * kTX * kTY * sizeof(float) <= 96 * 1024) ? 1 : 0;
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

template <typename T, bool FAST4, bool WANT_ARG>
static int launch_tail_fwd(const void* cost, const rsm_regress_out& out, const TailGeom& g, size_t smem, dim3 grid,
                           cudaStream_t st) {
  auto k = g.cached ? upsample_regress_fwd_kernel<T, FAST4, WANT_ARG, true>
                    : upsample_regress_fwd_kernel<T, FAST4, WANT_ARG, false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<grid, kTX * kTY, smem, st>>>((const T*)cost, (T*)out.soft, out.argmin, out.argmax, out.lse, g);
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
    if (g.fast4) return want_arg ? launch_tail_fwd<T, true, true>(cost, out, g, smem, grid, st)
                                 : launch_tail_fwd<T, true, false>(cost, out, g, smem, grid, st);
    return want_arg ? launch_tail_fwd<T, false, true>(cost, out, g, smem, grid, st)
                    : launch_tail_fwd<T, false, false>(cost, out, g, smem, grid, st);
  });
}

extern "C" int64_t rsm_upsample_regress_bwd_workspace(int64_t B, int64_t Dc, int64_t H, int64_t W) {
  if (B < 0 || Dc < 0 || H < 0 || W < 0) return -1;
  return B * Dc * H * W * (int64_t)sizeof(float);
}

extern "C" int rsm_upsample_regress_bwd(const void* gout, const void* cost, const void* soft,
                                        const float* lse, void* gcost, void* workspace, int64_t B,
                                        int64_t Dc, int64_t Hc, int64_t Wc, int64_t D, int64_t H,
                                        int64_t W, int dtype, int device, void* stream) {
  if (B < 0) return RSM_ERR_INVALID_SHAPE;
  TailGeom g;
  size_t smem;
  if (int rc = make_geom(Dc, Hc, Wc, D, H, W, g, smem)) return rc;
  if (B == 0) return RSM_OK;
  g.cached = 0;   // the adjoint evaluates every slice exactly once: nothing to park
  smem = tail_smem_bytes(g);
  if (!gout || !cost || !soft || !lse || !gcost || !workspace) return RSM_ERR_NULL_POINTER;
  if (B > 65535) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    auto k = upsample_regress_bwd_cols_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)B);
    if (grid.y > 65535) return (int)RSM_ERR_INVALID_SHAPE;
    k<<<grid, kTX * kTY, smem, st>>>((const T*)gout, (const T*)cost, (const T*)soft, lse, (float*)workspace, g);
    if (int rc = finish_launch("rsm_upsample_regress_bwd(cols)")) return rc;
    const int64_t total = B * Dc * Hc * Wc;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    upsample_regress_bwd_gather_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const float*)workspace, (T*)gcost, total, g);
    return finish_launch("rsm_upsample_regress_bwd(gather)");
  });
}
