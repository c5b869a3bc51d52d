// Disparity regression: soft-argmax (softmax(+cost) expectation) fused with hard argmin/argmax in
// one pass over a dense (N,D,H,W) cost, its adjoint, and the MobileStereoNetV4 head
// (trilinear upsample -> softmax -> expectation) evaluated without materialising (B,D,H,W).
#include <math.h>

#include "rsm_common.cuh"

namespace rsm {

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
  __device__ __forceinline__ void update(float v, int d) {
    const bool vnan = v != v;
    if ((v < minv) || (vnan && minv == minv)) { minv = v; mini = d; }
    if ((v > maxv) || (vnan && maxv == maxv)) { maxv = v; maxi = d; }
  }
};

// ============================================================================ regress fwd
// One thread owns VEC consecutive pixels and streams the D values of each (coalesced 16-byte
// loads across the warp, stride H*W between disparities), DCH disparities per step so that one
// rescale exp serves DCH value exps and DCH independent loads are in flight.
template <typename T, int VEC, int DCH>
__global__ void __launch_bounds__(256)
regress_fwd_kernel(const T* __restrict__ cost, T* __restrict__ soft, int64_t* __restrict__ amin,
                   int64_t* __restrict__ amax, float* __restrict__ lse, int64_t pix_vec_per_img,
                   int64_t total_vec, int D, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t n = i / pix_vec_per_img;
  const int64_t p = (i - n * pix_vec_per_img) * VEC;
  const T* __restrict__ base = cost + n * D * HW + p;

  float m[VEC], s[VEC], ws[VEC];
  ArgTrack trk[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { m[j] = -INFINITY; s[j] = 0.f; ws[j] = 0.f; }

  for (int d0 = 0; d0 < D; d0 += DCH) {
    float v[DCH][VEC];
#pragma unroll
    for (int k = 0; k < DCH; ++k) {
      if (d0 + k < D) {
        if constexpr (VEC * sizeof(T) == 16) {
          Vec16<T> t = ldcs16(base + (int64_t)(d0 + k) * HW);
#pragma unroll
          for (int j = 0; j < VEC; ++j) v[k][j] = to_f(t.v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) v[k][j] = to_f(__ldcs(base + (int64_t)(d0 + k) * HW + j));
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[k][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float cm = v[0][j];
#pragma unroll
      for (int k = 1; k < DCH; ++k) cm = fmaxf(cm, v[k][j]);
      const float mn = fmaxf(m[j], cm);
      const float mnl = mn * kLog2e;
      // exp(m - mn); when both are -inf (nothing seen yet) the scale is irrelevant: s = ws = 0
      const float a = (m[j] == -INFINITY) ? 0.f : fast_exp2(m[j] * kLog2e - mnl);
      float sj = s[j] * a, wj = ws[j] * a;
#pragma unroll
      for (int k = 0; k < DCH; ++k) {
        if (d0 + k < D) {
          const float e = fast_exp2(fmaf(v[k][j], kLog2e, -mnl));
          sj += e;
          wj = fmaf((float)(d0 + k), e, wj);
          trk[j].update(v[k][j], d0 + k);
        }
      }
      m[j] = mn; s[j] = sj; ws[j] = wj;
    }
  }
  const int64_t o = n * HW + p;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    if (soft) soft[o + j] = from_f<T>(ws[j] / s[j]);
    if (lse) lse[o + j] = m[j] + __logf(s[j]);
    if (amin) amin[o + j] = trk[j].mini;
    if (amax) amax[o + j] = trk[j].maxi;
  }
}

// ============================================================================ regress bwd
// gcost[d] = g * softmax_d * (d - E), with softmax_d = exp(cost[d] - lse)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
regress_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ cost, const T* __restrict__ soft,
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
    e[j] = to_f(soft[o + j]);
    l2[j] = lse[o + j] * kLog2e;
  }
  const T* __restrict__ cb = cost + n * D * HW + p;
  T* __restrict__ gb = gcost + n * D * HW + p;
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
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

// =============================================================== v4 head: upsample + regress
// Linear-interpolation source index/weights of F.interpolate(align_corners=False) as ATen
// computes them in fp32 (no FMA contraction so the host-side oracle table matches bit for bit).
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

constexpr int kTX = 32, kTY = 8;  // fine-pixel tile of one CTA

struct TailGeom {
  int Dc, Hc, Wc, D, H, W;
  int FH, FW;                    // footprint (coarse rows / cols) a tile can touch, upper bound
  float sd, sh, sw;              // in/out scales per axis
};

// stage the coarse footprint of this tile: sm[k][fy][fx] (fp32), rows cy0.., cols cx0..
template <typename T>
__device__ __forceinline__ void stage_footprint(const T* __restrict__ cost_b, float* sm, const TailGeom& g,
                                                int cy0, int cx0) {
  const int per = g.FH * g.FW;
  const int tot = g.Dc * per;
  for (int e = threadIdx.x; e < tot; e += kTX * kTY) {
    const int k = e / per;
    const int r = e - k * per;
    const int fy = r / g.FW, fx = r - fy * g.FW;
    const int cy = min(cy0 + fy, g.Hc - 1), cx = min(cx0 + fx, g.Wc - 1);
    sm[e] = to_f(__ldg(cost_b + ((int64_t)k * g.Hc + cy) * g.Wc + cx));
  }
}

template <typename T>
__global__ void __launch_bounds__(kTX * kTY)
upsample_regress_fwd_kernel(const T* __restrict__ cost, T* __restrict__ soft, int64_t* __restrict__ amin,
                            int64_t* __restrict__ amax, float* __restrict__ lse, TailGeom g) {
  extern __shared__ float sm[];
  const int b = blockIdx.z;
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x = blockIdx.x * kTX + tx, y = blockIdx.y * kTY + ty;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_footprint(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  __syncthreads();
  if (x >= g.W || y >= g.H) return;

  const Lin ly = lin_index(y, g.sh, g.Hc), lx = lin_index(x, g.sw, g.Wc);
  const int per = g.FH * g.FW;
  const int o00 = (ly.i0 - cy0) * g.FW + (lx.i0 - cx0), o01 = (ly.i0 - cy0) * g.FW + (lx.i1 - cx0);
  const int o10 = (ly.i1 - cy0) * g.FW + (lx.i0 - cx0), o11 = (ly.i1 - cy0) * g.FW + (lx.i1 - cx0);
  auto slice = [&](int k) -> float {
    const float* s = sm + k * per;
    return ly.w0 * (lx.w0 * s[o00] + lx.w1 * s[o01]) + ly.w1 * (lx.w0 * s[o10] + lx.w1 * s[o11]);
  };
  // pass 1: stabiliser.  Fine values are convex combinations of the coarse slices, so the max
  // over the coarse slices bounds them (and is attained within |delta|/8 for the x4 head).
  float M = -INFINITY;
  for (int k = 0; k < g.Dc; ++k) M = fmaxf(M, slice(k));
  const float Ml = M * kLog2e;
  // pass 2: stream the fine disparities in ascending order; (i0, i1) only move forward
  float s = 0.f, ws = 0.f;
  ArgTrack trk;
  int k0 = -1, k1 = -1;
  float c0 = 0.f, c1 = 0.f;
  for (int d = 0; d < g.D; ++d) {
    const Lin ld = lin_index(d, g.sd, g.Dc);
    if (ld.i0 != k0) { c0 = (ld.i0 == k1) ? c1 : slice(ld.i0); k0 = ld.i0; }
    if (ld.i1 != k1) { c1 = (ld.i1 == k0) ? c0 : slice(ld.i1); k1 = ld.i1; }
    const float f = ld.w0 * c0 + ld.w1 * c1;
    const float e = fast_exp2(fmaf(f, kLog2e, -Ml));
    s += e;
    ws = fmaf((float)d, e, ws);
    trk.update(f, d);
  }
  const int64_t o = ((int64_t)b * g.H + y) * g.W + x;
  if (soft) soft[o] = from_f<T>(ws / s);
  if (lse) lse[o] = M + __logf(s);
  if (amin) amin[o] = trk.mini;
  if (amax) amax[o] = trk.maxi;
}

// backward, stage 1: per fine pixel, gradient with respect to the bilinearly interpolated coarse
// column c_k at that pixel -> ws (B,Dc,H,W) fp32.  Deterministic (no atomics).
template <typename T>
__global__ void __launch_bounds__(kTX * kTY)
upsample_regress_bwd_cols_kernel(const T* __restrict__ gout, const T* __restrict__ cost,
                                 const T* __restrict__ soft, const float* __restrict__ lse,
                                 float* __restrict__ wsp, TailGeom g) {
  extern __shared__ float sm[];
  const int b = blockIdx.z;
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x = blockIdx.x * kTX + tx, y = blockIdx.y * kTY + ty;
  const int cy0 = lin_index(blockIdx.y * kTY, g.sh, g.Hc).i0;
  const int cx0 = lin_index(blockIdx.x * kTX, g.sw, g.Wc).i0;
  stage_footprint(cost + (int64_t)b * g.Dc * g.Hc * g.Wc, sm, g, cy0, cx0);
  __syncthreads();
  if (x >= g.W || y >= g.H) return;

  const Lin ly = lin_index(y, g.sh, g.Hc), lx = lin_index(x, g.sw, g.Wc);
  const int per = g.FH * g.FW;
  const int o00 = (ly.i0 - cy0) * g.FW + (lx.i0 - cx0), o01 = (ly.i0 - cy0) * g.FW + (lx.i1 - cx0);
  const int o10 = (ly.i1 - cy0) * g.FW + (lx.i0 - cx0), o11 = (ly.i1 - cy0) * g.FW + (lx.i1 - cx0);
  auto slice = [&](int k) -> float {
    const float* s = sm + k * per;
    return ly.w0 * (lx.w0 * s[o00] + lx.w1 * s[o01]) + ly.w1 * (lx.w0 * s[o10] + lx.w1 * s[o11]);
  };
  const int64_t o = ((int64_t)b * g.H + y) * g.W + x;
  const float go = to_f(gout[o]), E = to_f(soft[o]), l2 = lse[o] * kLog2e;
  const int64_t plane = (int64_t)g.H * g.W;
  float* __restrict__ col = wsp + (int64_t)b * g.Dc * plane + (int64_t)y * g.W + x;

  int cur = 0;               // coarse slice accA belongs to; accB belongs to cur + 1
  float accA = 0.f, accB = 0.f;
  int k0 = -1, k1 = -1;
  float c0 = 0.f, c1 = 0.f;
  for (int d = 0; d < g.D; ++d) {
    const Lin ld = lin_index(d, g.sd, g.Dc);
    while (cur < ld.i0) {    // slices below i0 are complete
      col[(int64_t)cur * plane] = accA;
      accA = accB; accB = 0.f; ++cur;
    }
    if (ld.i0 != k0) { c0 = (ld.i0 == k1) ? c1 : slice(ld.i0); k0 = ld.i0; }
    if (ld.i1 != k1) { c1 = (ld.i1 == k0) ? c0 : slice(ld.i1); k1 = ld.i1; }
    const float f = ld.w0 * c0 + ld.w1 * c1;
    const float p = fast_exp2(fmaf(f, kLog2e, -l2));
    const float gf = go * p * ((float)d - E);
    accA = fmaf(ld.w0, gf, accA);
    if (ld.i1 == ld.i0) accA = fmaf(ld.w1, gf, accA); else accB = fmaf(ld.w1, gf, accB);
  }
  col[(int64_t)cur * plane] = accA;
  if (cur + 1 < g.Dc) col[(int64_t)(cur + 1) * plane] = accB;
  for (int k = cur + 2; k < g.Dc; ++k) col[(int64_t)k * plane] = 0.f;
}

// range of fine indices o whose (i0 or i1) can equal coarse index ic (conservative; exact test inside)
__device__ __forceinline__ void fine_range(int ic, float scale, int n_out, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = max(0, (int)floorf(((float)ic - 0.5f) * inv - 0.5f) - 1);
  hi = min(n_out - 1, (int)ceilf(((float)ic + 1.5f) * inv - 0.5f) + 1);
}

// backward, stage 2: transposed bilinear gather, one thread per coarse element
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
    const float wy = (ly.i0 == yc ? ly.w0 : 0.f) + (ly.i1 == yc ? ly.w1 : 0.f);
    if (wy == 0.f && ly.i0 != yc && ly.i1 != yc) continue;
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
  smem = (size_t)g.Dc * g.FH * g.FW * sizeof(float);
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

extern "C" int rsm_regress_fwd(const void* cost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                               rsm_regress_out out, int device, void* stream) {
  if (N < 0 || D <= 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;  // softmax over an empty axis is undefined
  if (N * H * W == 0) return RSM_OK;
  if (!cost) return RSM_ERR_NULL_POINTER;
  if (D > 2147483647LL) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t HW = H * W;
    const bool vec = HW % VEC == 0 && aligned_to(cost, 16);
    const int64_t pv = vec ? HW / VEC : HW;
    const int64_t total = N * pv;
    if (!grid_ok(ceil_div(total, 256))) return (int)RSM_ERR_INVALID_SHAPE;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    if (vec)
      regress_fwd_kernel<T, VEC, (sizeof(T) == 4 ? 8 : 4)><<<blocks, 256, 0, st>>>(
          (const T*)cost, (T*)out.soft, out.argmin, out.argmax, out.lse, pv, total, (int)D, HW);
    else
      regress_fwd_kernel<T, 1, 8><<<blocks, 256, 0, st>>>((const T*)cost, (T*)out.soft, out.argmin, out.argmax,
                                                          out.lse, pv, total, (int)D, HW);
    return finish_launch("rsm_regress_fwd");
  });
}

extern "C" int rsm_regress_bwd(const void* gout, const void* cost, const void* soft, const float* lse,
                               void* gcost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                               int device, void* stream) {
  if (N < 0 || D <= 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * H * W == 0) return RSM_OK;
  if (!gout || !cost || !soft || !lse || !gcost) return RSM_ERR_NULL_POINTER;
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
      regress_bwd_kernel<T, VEC><<<blocks, 256, 0, st>>>((const T*)gout, (const T*)cost, (const T*)soft, lse,
                                                         (T*)gcost, pv, total, (int)D, HW);
    else
      regress_bwd_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)gout, (const T*)cost, (const T*)soft, lse,
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
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    auto k = upsample_regress_fwd_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)B);
    if (grid.y > 65535) return (int)RSM_ERR_INVALID_SHAPE;
    k<<<grid, kTX * kTY, smem, st>>>((const T*)cost, (T*)out.soft, out.argmin, out.argmax, out.lse, g);
    return finish_launch("rsm_upsample_regress_fwd");
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
