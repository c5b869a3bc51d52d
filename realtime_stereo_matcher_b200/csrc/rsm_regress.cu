// Disparity regression: soft-argmax (softmax(+cost) expectation) fused with hard argmin/argmax in
// one pass over a dense (N,D,H,W) cost, its adjoint, and the expectation of given probabilities.
// (The MobileStereoNetV4 head lives in rsm_tail.cu.)
#include <math.h>

#include "rsm_common.cuh"

namespace rsm {

// ============================================================================ regress fwd
// One thread owns VEC consecutive pixels and streams their D values (coalesced vector loads across
// the warp, stride H*W between disparities), DCH disparities per step: DCH independent loads in
// flight, one rescale exp per DCH value exps (chunked online softmax).  The arg-extrema follow
// torch.argmin/argmax: first index on ties (strict compares in ascending d) and NaN as the
// extremum -- NaNs only raise a flag in the hot loop; flagged pixels (rare) are re-scanned for
// their first NaN at the end.
template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC * sizeof(T) == 16) {
    Vec16<T> t = ldcs16(p);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = to_f(t.v[j]);
  } else if constexpr (VEC * sizeof(T) == 8) {
    union { uint2 raw; T e[VEC]; } t;
    t.raw = __ldcs(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = to_f(t.e[j]);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = to_f(__ldcs(p + j));
  }
}

template <int VEC, bool SOFT, bool ARG>
struct RegressState {
  float m[VEC], s[VEC], ws[VEC];
  float minv[VEC], maxv[VEC];
  int mini[VEC], maxi[VEC];
  bool nan[VEC];
  __device__ __forceinline__ RegressState() {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      m[j] = -INFINITY; s[j] = 0.f; ws[j] = 0.f;
      minv[j] = INFINITY; maxv[j] = -INFINITY; mini[j] = 0; maxi[j] = 0; nan[j] = false;
    }
  }
  // a full chunk of 8 disparities d0 .. d0+7, lean form (the combined soft + arg kernel is issue-bound): extrema
  // by FMNMX trees, their first index by equality selects, ONE running-extremum update per chunk, compile-time
  // exponent weights on top of a per-chunk float base.  The NaN flag is conservative (sum of exponentials, or sum
  // of the values for the arg-only kernel, is NaN): the rescan at the end decides.  Same torch semantics.
  __device__ __forceinline__ void consume_full8(const float (&v)[8][VEC], int d0) {
    const float fb = (float)d0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float cmax = fmaxf(fmaxf(fmaxf(v[0][j], v[1][j]), fmaxf(v[2][j], v[3][j])),
                               fmaxf(fmaxf(v[4][j], v[5][j]), fmaxf(v[6][j], v[7][j])));
      if constexpr (SOFT) {
        const float mn = fmaxf(m[j], cmax);
        const float mnl = mn * kLog2e;
        const float a = (m[j] == -INFINITY) ? 0.f : fast_exp2(fmaf(m[j], kLog2e, -mnl));
        float e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = fast_exp2(fmaf(v[k][j], kLog2e, -mnl));
        const float S = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
        const float T = (fmaf(2.f, e[2], e[1]) + fmaf(3.f, e[3], 4.f * e[4])) + (fmaf(5.f, e[5], 6.f * e[6]) + 7.f * e[7]);
        s[j] = fmaf(s[j], a, S);
        ws[j] = fmaf(ws[j], a, fmaf(fb, S, T));
        m[j] = mn;
        if constexpr (ARG) nan[j] |= (S != S);
      }
      if constexpr (ARG) {
        const float cmin = fminf(fminf(fminf(v[0][j], v[1][j]), fminf(v[2][j], v[3][j])),
                                 fminf(fminf(v[4][j], v[5][j]), fminf(v[6][j], v[7][j])));
        int imin = 7, imax = 7;
#pragma unroll
        for (int k = 6; k >= 0; --k) {
          imin = v[k][j] == cmin ? k : imin;
          imax = v[k][j] == cmax ? k : imax;
        }
        if (cmin < minv[j]) { minv[j] = cmin; mini[j] = d0 + imin; }
        if (cmax > maxv[j]) { maxv[j] = cmax; maxi[j] = d0 + imax; }
        if constexpr (!SOFT) {
          const float vs = ((v[0][j] + v[1][j]) + (v[2][j] + v[3][j])) + ((v[4][j] + v[5][j]) + (v[6][j] + v[7][j]));
          nan[j] |= (vs != vs);
        }
      }
    }
  }
  // consume disparities d0 .. d0+CNT-1 held in v[k][j]
  template <int DCH>
  __device__ __forceinline__ void consume(const float (&v)[DCH][VEC], int d0, int cnt) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if constexpr (SOFT) {
        float cm = v[0][j];
#pragma unroll
        for (int k = 1; k < DCH; ++k)
          if (k < cnt) cm = fmaxf(cm, v[k][j]);
        const float mn = fmaxf(m[j], cm);
        const float mnl = mn * kLog2e;
        // exp(m - mn); nothing seen yet (m = -inf): s = ws = 0 and the scale is irrelevant
        const float a = (m[j] == -INFINITY) ? 0.f : fast_exp2(fmaf(m[j], kLog2e, -mnl));
        float sj = s[j] * a, wj = ws[j] * a;
#pragma unroll
        for (int k = 0; k < DCH; ++k)
          if (k < cnt) {
            const float e = fast_exp2(fmaf(v[k][j], kLog2e, -mnl));
            sj += e;
            wj = fmaf((float)(d0 + k), e, wj);
          }
        m[j] = mn; s[j] = sj; ws[j] = wj;
      }
      if constexpr (ARG) {
#pragma unroll
        for (int k = 0; k < DCH; ++k)
          if (k < cnt) {
            const float x = v[k][j];
            if (x < minv[j]) { minv[j] = x; mini[j] = d0 + k; }
            if (x > maxv[j]) { maxv[j] = x; maxi[j] = d0 + k; }
            nan[j] |= (x != x);
          }
      }
    }
  }
};

template <typename T, int VEC, int DCH, bool SOFT, bool ARG>
__global__ void __launch_bounds__(256)
regress_fwd_kernel(const T* __restrict__ cost, T* __restrict__ soft, int64_t* __restrict__ amin,
                   int64_t* __restrict__ amax, float* __restrict__ lse, float* __restrict__ expect, int64_t pix_vec_per_img,
                   int64_t total_vec, int D, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t n = i / pix_vec_per_img;
  const int64_t p = (i - n * pix_vec_per_img) * VEC;
  const T* __restrict__ base = cost + n * D * HW + p;

  RegressState<VEC, SOFT, ARG> st;
  int d0 = 0;
  for (; d0 + DCH <= D; d0 += DCH) {      // full chunks: no bounds checks
    float v[DCH][VEC];
#pragma unroll
    for (int k = 0; k < DCH; ++k) load_vec<T, VEC>(base + (int64_t)(d0 + k) * HW, v[k]);
    if constexpr (DCH == 8) st.consume_full8(v, d0);
    else st.template consume<DCH>(v, d0, DCH);
  }
  if (d0 < D) {                            // ragged tail
    float v[DCH][VEC];
#pragma unroll
    for (int k = 0; k < DCH; ++k) {
      if (d0 + k < D) load_vec<T, VEC>(base + (int64_t)(d0 + k) * HW, v[k]);
      else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[k][j] = 0.f;
      }
    }
    st.template consume<DCH>(v, d0, D - d0);
  }
  const int64_t o = n * HW + p;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    if constexpr (SOFT) {
      const float e = st.ws[j] / st.s[j];
      if (soft) soft[o + j] = from_f<T>(e);
      if (expect) expect[o + j] = e;
      if (lse) lse[o + j] = st.m[j] + __logf(st.s[j]);
    }
    if constexpr (ARG) {
      int mi = st.mini[j], ma = st.maxi[j];
      if (st.nan[j]) {                     // torch: NaN is both the minimum and the maximum; first one wins
        for (int d = 0; d < D; ++d) {
          const float x = to_f(base[(int64_t)d * HW + j]);
          if (x != x) { mi = ma = d; break; }
        }
      }
      if (amin) amin[o + j] = mi;
      if (amax) amax[o + j] = ma;
    }
  }
}

// ============================================================================ regress bwd
// gcost[d] = g * softmax_d * (d - E), with softmax_d = exp(cost[d] - lse)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
regress_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ cost, const float* __restrict__ expect,
                   const float* __restrict__ lse, T* __restrict__ gcost, int64_t pix_vec_per_img,
                   int64_t total_vec, int D, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t n = i / pix_vec_per_img;
  const int64_t p = (i - n * pix_vec_per_img) * VEC;
  const int64_t o = n * HW + p;
  float g[VEC], e[VEC], l2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    g[j] = to_f(gout[o + j]);
    e[j] = expect[o + j];
    l2[j] = lse[o + j] * kLog2e;
  }
  const T* __restrict__ cb = cost + n * D * HW + p;
  T* __restrict__ gb = gcost + n * D * HW + p;
  int d = 0;
  if constexpr (VEC * sizeof(T) == 16) {
    // eight independent 128-bit loads in flight per thread, like the forward pass
    constexpr int U = 8;
    for (; d + U <= D; d += U) {
      Vec16<T> t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) t[u] = ldcs16(cb + (int64_t)(d + u) * HW);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        Vec16<T> r;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float pr = fast_exp2(fmaf(to_f(t[u].v[j]), kLog2e, -l2[j]));
          r.v[j] = from_f<T>(g[j] * pr * ((float)(d + u) - e[j]));
        }
        stcs16(gb + (int64_t)(d + u) * HW, r);
      }
    }
  }
#pragma unroll 4
  for (; d < D; ++d) {
    if constexpr (VEC * sizeof(T) == 16) {
      Vec16<T> t = ldcs16(cb + (int64_t)d * HW), r;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float pr = fast_exp2(fmaf(to_f(t.v[j]), kLog2e, -l2[j]));
        r.v[j] = from_f<T>(g[j] * pr * ((float)d - e[j]));
      }
      stcs16(gb + (int64_t)d * HW, r);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float pr = fast_exp2(fmaf(to_f(cb[(int64_t)d * HW + j]), kLog2e, -l2[j]));
        gb[(int64_t)d * HW + j] = from_f<T>(g[j] * pr * ((float)d - e[j]));
      }
    }
  }
}

// ============================================================================== expectation
// disparity_regression of probabilities (v4 form): out = sum_d d * prob[d]; adjoint gprob[d] = g * d
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
expect_fwd_kernel(const T* __restrict__ prob, T* __restrict__ out, int64_t pix_vec_per_img, int64_t total_vec,
                  int D, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t n = i / pix_vec_per_img;
  const int64_t p = (i - n * pix_vec_per_img) * VEC;
  const T* __restrict__ base = prob + n * D * HW + p;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < D; ++d) {
    if constexpr (VEC * sizeof(T) == 16) {
      Vec16<T> t = ldcs16(base + (int64_t)d * HW);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = fmaf((float)d, to_f(t.v[j]), acc[j]);
    } else {
      acc[0] = fmaf((float)d, to_f(__ldcs(base + (int64_t)d * HW)), acc[0]);
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) out[n * HW + p + j] = from_f<T>(acc[j]);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
expect_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gprob, int64_t pix_vec_per_img, int64_t total_vec,
                  int D, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t n = i / pix_vec_per_img;
  const int64_t p = (i - n * pix_vec_per_img) * VEC;
  float g[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) g[j] = to_f(gout[n * HW + p + j]);
  T* __restrict__ base = gprob + n * D * HW + p;
  for (int d = 0; d < D; ++d) {
    if constexpr (VEC * sizeof(T) == 16) {
      Vec16<T> r;
#pragma unroll
      for (int j = 0; j < VEC; ++j) r.v[j] = from_f<T>(g[j] * (float)d);
      stcs16(base + (int64_t)d * HW, r);
    } else {
      base[(int64_t)d * HW] = from_f<T>(g[0] * (float)d);
    }
  }
}

static bool grid_ok(int64_t blocks) { return blocks >= 0 && blocks <= 2147483647LL; }

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

extern "C" int rsm_regress_fwd(const void* cost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                               rsm_regress_out out, int device, void* stream) {
  if (N < 0 || D <= 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;  // softmax over an empty axis is undefined
  if (N * H * W == 0) return RSM_OK;
  if (!cost) return RSM_ERR_NULL_POINTER;
  if (D > 2147483647LL) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 4;   // 16-byte loads for fp32, 8-byte loads for 16-bit: same register budget
    const int64_t HW = H * W;
    const bool vec = HW % VEC == 0 && aligned_to(cost, VEC * sizeof(T));
    const int64_t pv = vec ? HW / VEC : HW;
    const int64_t total = N * pv;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    const bool want_soft = out.soft || out.lse || out.expect, want_arg = out.argmin || out.argmax;
    auto launch = [&](auto kern) {
      kern<<<blocks, 256, 0, st>>>((const T*)cost, (T*)out.soft, out.argmin, out.argmax, out.lse, out.expect, pv, total, (int)D, HW);
    };
    if (vec) {
      if (want_soft && want_arg) launch(regress_fwd_kernel<T, 4, 8, true, true>);
      else if (want_soft) launch(regress_fwd_kernel<T, 4, 8, true, false>);
      else launch(regress_fwd_kernel<T, 4, 8, false, true>);
    } else {
      if (want_soft && want_arg) launch(regress_fwd_kernel<T, 1, 8, true, true>);
      else if (want_soft) launch(regress_fwd_kernel<T, 1, 8, true, false>);
      else launch(regress_fwd_kernel<T, 1, 8, false, true>);
    }
    return finish_launch("rsm_regress_fwd");
  });
}

extern "C" int rsm_regress_bwd(const void* gout, const void* cost, const float* expect, const float* lse,
                               void* gcost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                               int device, void* stream) {
  if (N < 0 || D <= 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * H * W == 0) return RSM_OK;
  if (!gout || !cost || !expect || !lse || !gcost) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t HW = H * W;
    const bool vec = HW % VEC == 0 && aligned_to(cost, 16) && aligned_to(gcost, 16);
    const int64_t pv = vec ? HW / VEC : HW;
    const int64_t total = N * pv;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    if (vec)
      regress_bwd_kernel<T, VEC><<<blocks, 256, 0, st>>>((const T*)gout, (const T*)cost, expect, lse,
                                                         (T*)gcost, pv, total, (int)D, HW);
    else
      regress_bwd_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)gout, (const T*)cost, expect, lse,
                                                       (T*)gcost, pv, total, (int)D, HW);
    return finish_launch("rsm_regress_bwd");
  });
}

extern "C" int rsm_expect_fwd(const void* prob, void* out, int64_t N, int64_t D, int64_t H, int64_t W,
                              int dtype, int device, void* stream) {
  if (N < 0 || D < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * H * W == 0) return RSM_OK;
  if (!out || (D > 0 && !prob)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t HW = H * W;
    const bool vec = HW % VEC == 0 && aligned_to(prob, 16);
    const int64_t pv = vec ? HW / VEC : HW;
    const int64_t total = N * pv;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    if (vec) expect_fwd_kernel<T, VEC><<<blocks, 256, 0, st>>>((const T*)prob, (T*)out, pv, total, (int)D, HW);
    else expect_fwd_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)prob, (T*)out, pv, total, (int)D, HW);
    return finish_launch("rsm_expect_fwd");
  });
}

extern "C" int rsm_expect_bwd(const void* gout, void* gprob, int64_t N, int64_t D, int64_t H, int64_t W,
                              int dtype, int device, void* stream) {
  if (N < 0 || D < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * H * W * D == 0) return RSM_OK;
  if (!gout || !gprob) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t HW = H * W;
    const bool vec = HW % VEC == 0 && aligned_to(gprob, 16);
    const int64_t pv = vec ? HW / VEC : HW;
    const int64_t total = N * pv;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    if (vec) expect_bwd_kernel<T, VEC><<<blocks, 256, 0, st>>>((const T*)gout, (T*)gprob, pv, total, (int)D, HW);
    else expect_bwd_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)gout, (T*)gprob, pv, total, (int)D, HW);
    return finish_launch("rsm_expect_bwd");
  });
}

