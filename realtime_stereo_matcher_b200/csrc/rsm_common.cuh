// Shared device/host helpers for librsm_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rsm.h"

namespace rsm {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ----------------------------------------------------------------------------- host side
// thread-local detail of the last CUDA failure (read by rsm_last_error)
void set_cuda_error(cudaError_t e, const char* where);

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    want = device;
  }
  ~DeviceGuard() {
    if (ok && prev != want) cudaSetDevice(prev);
  }
  int want = -1;
};

inline int finish_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e, where);
    return RSM_ERR_CUDA;
  }
  return RSM_OK;
}

inline bool valid_dtype(int dt) { return dt == RSM_F32 || dt == RSM_F16 || dt == RSM_BF16; }
inline int dtype_size(int dt) { return dt == RSM_F32 ? 4 : 2; }
inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// dispatch a generic lambda on a runtime dtype: f(T{}) with T in {float, __half, __nv_bfloat16}
#define RSM_DISPATCH_DTYPE(dt, T, ...)                           \
  [&]() -> int {                                                 \
    switch (dt) {                                                \
      case RSM_F32: { using T = float; return __VA_ARGS__(); }   \
      case RSM_F16: { using T = __half; return __VA_ARGS__(); }  \
      case RSM_BF16: { using T = __nv_bfloat16; return __VA_ARGS__(); } \
      default: return (int)RSM_ERR_UNSUPPORTED_DTYPE;            \
    }                                                            \
  }()

// ---------------------------------------------------------------------------- device side
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T v[N];
  };
};

template <typename T> __device__ __forceinline__ Vec16<T> ldg16(const T* p) {
  Vec16<T> r;
  r.raw = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
// streaming (evict-first) 16-byte load / store for data touched exactly once
template <typename T> __device__ __forceinline__ Vec16<T> ldcs16(const T* p) {
  Vec16<T> r;
  r.raw = __ldcs(reinterpret_cast<const uint4*>(p));
  return r;
}
template <typename T> __device__ __forceinline__ void stcs16(T* p, const Vec16<T>& v) {
  __stcs(reinterpret_cast<uint4*>(p), v.raw);
}
template <typename T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& v) {
  *reinterpret_cast<uint4*>(p) = v.raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// running first-occurrence argmin / argmax with torch semantics (NaN is the extremum)
struct ArgTrack {
  float minv = INFINITY, maxv = -INFINITY;
  int mini = 0, maxi = 0;
  __device__ __forceinline__ void update_min(float v, int d) {
    if ((v < minv) || (v != v && minv == minv)) { minv = v; mini = d; }
  }
  __device__ __forceinline__ void update_max(float v, int d) {
    if ((v > maxv) || (v != v && maxv == maxv)) { maxv = v; maxi = d; }
  }
  __device__ __forceinline__ void update(float v, int d) { update_min(v, d); update_max(v, d); }
};

// Per-pixel running soft-argmax / arg-extrema state, fed 8 disparities at a time in ascending order (torch
// semantics: first index wins ties; NaN wins, first one).  Lean form for issue-bound epilogues: extrema by FMNMX
// trees, their first index by equality selects, one running-extremum update per chunk, compile-time exponent
// weights on top of a per-chunk float base, NaNs looked for only when the chunk's sum of exponentials is NaN.
struct ScanState {
  float m = -INFINITY, s = 0.f, ws = 0.f, minv = INFINITY, maxv = -INFINITY;
  int mini = 0, maxi = 0, nani = 0x7fffffff;
  __device__ __forceinline__ void chunk8(const float (&v)[8], int d0) {
    const float cmin = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(v[4], v[5]), fminf(v[6], v[7])));
    const float cmax = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
    int imin = 7, imax = 7;
#pragma unroll
    for (int k = 6; k >= 0; --k) {
      imin = v[k] == cmin ? k : imin;
      imax = v[k] == cmax ? k : imax;
    }
    if (cmin < minv) { minv = cmin; mini = d0 + imin; }
    if (cmax > maxv) { maxv = cmax; maxi = d0 + imax; }
    const float mn = fmaxf(m, cmax), mnl = mn * kLog2e;
    const float a = (m == -INFINITY) ? 0.f : fast_exp2(fmaf(m, kLog2e, -mnl));
    float e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fast_exp2(fmaf(v[k], kLog2e, -mnl));
    const float S = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    const float T = (fmaf(2.f, e[2], e[1]) + fmaf(3.f, e[3], 4.f * e[4])) + (fmaf(5.f, e[5], 6.f * e[6]) + 7.f * e[7]);
    s = fmaf(s, a, S);
    ws = fmaf(ws, a, fmaf((float)d0, S, T));
    m = mn;
    if (S != S) {
#pragma unroll
      for (int k = 7; k >= 0; --k)
        if (v[k] != v[k]) nani = min(nani, d0 + k);
    }
  }
  // ragged tail: cnt < 8 values
  __device__ __forceinline__ void tail(const float (&v)[8], int d0, int cnt) {
    float gm = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < cnt) {
        const float f = v[k];
        if (f < minv) { minv = f; mini = d0 + k; }
        if (f > maxv) { maxv = f; maxi = d0 + k; }
        if (f != f) nani = min(nani, d0 + k);
        gm = fmaxf(gm, f);
      }
    const float mn = fmaxf(m, gm), mnl = mn * kLog2e;
    const float a = (m == -INFINITY) ? 0.f : fast_exp2(fmaf(m, kLog2e, -mnl));
    s *= a; ws *= a;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < cnt) {
        const float e = fast_exp2(fmaf(v[k], kLog2e, -mnl));
        s += e;
        ws = fmaf((float)(d0 + k), e, ws);
      }
    m = mn;
  }
  __device__ __forceinline__ void finish() {
    if (nani != 0x7fffffff) { mini = nani; maxi = nani; }
  }
};

struct FeatView {  // device-side copy of rsm_feat
  const void* data;
  int64_t sn, sc, sh, sw;
};
inline FeatView view_of(const rsm_feat& f) { return FeatView{f.data, f.stride_n, f.stride_c, f.stride_h, f.stride_w}; }

}  // namespace rsm
