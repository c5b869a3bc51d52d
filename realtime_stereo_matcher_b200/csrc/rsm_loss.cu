// Loss and metrics on the device: SequenceLoss.forward (loss/loss.py:36-81) and
// get_flow_map_metrics (loss/loss.py:6-22) as masked reductions, forward and adjoint.
//
// One prediction of the sequence per call: valid = (flow_valid >= 0.5) & (|flow_gt| < max_flow)
// (loss.py:55-58, one-channel flow), the prediction is brought to the ground-truth size the way the reference does
// (F.interpolate(pred * scale, size), nearest, loss.py:71-73 -- ATen's index rule), the element loss is
// L1 (|gt - p|) or smooth-L1 with beta = 1 (the last prediction, loss.py:75-78), and the result is the mean over
// the valid pixels.  The reference materialises the resized prediction, the loss map and the boolean gather;
// here everything is one read of pred / gt / valid.  Sums are carried in fp64 through a fixed-shape two-stage
// reduction (at most RSM_REDUCE_MAX_BLOCKS partials, combined in index order): deterministic.
// The non-finite checks of loss.py:60,66-67 are counted on the way and returned with the result, so that the caller
// can test them with ONE device read instead of a synchronisation per assert.
// Metrics: end-point error over the valid pixels (mean, fractions under 0.5 / 1 / 3 / 5 px) and min / max of the
// first batch item, in one pass, one device read instead of seven .item() calls (train_stereo.py:174).
// HBM-bound; algorithmic bytes = pred + gt + valid read once.
#include <algorithm>

#include "rsm_common.cuh"

namespace rsm {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = RSM_REDUCE_MAX_BLOCKS;   // 8 per SM
constexpr int RED_SLOTS = 8;                            // doubles per partial

// torch.min / torch.max propagate NaN (loss/loss.py:19-20 logs them as the divergence tell-tale, train_stereo.py:174);
// fmin / fmax would drop it
__device__ __forceinline__ double nan_min(double a, double b) { return (a != a || b != b) ? (double)NAN : fmin(a, b); }
__device__ __forceinline__ double nan_max(double a, double b) { return (a != a || b != b) ? (double)NAN : fmax(a, b); }
__device__ __forceinline__ float nan_minf(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }
__device__ __forceinline__ float nan_maxf(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], const bool (&is_min)[K], const bool (&is_max)[K],
                                                   double* __restrict__ dst) {
  __shared__ double sh[RED_THREADS / 32][K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_down_sync(0xffffffffu, v[k], o);
      v[k] = is_min[k] ? nan_min(v[k], other) : is_max[k] ? nan_max(v[k], other) : v[k] + other;
    }
    if (lane == 0) sh[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double a = sh[0][k];
      for (int w = 1; w < RED_THREADS / 32; ++w)
        a = is_min[k] ? nan_min(a, sh[w][k]) : is_max[k] ? nan_max(a, sh[w][k]) : a + sh[w][k];
      dst[k] = a;
    }
  }
}

struct LossGeom {
  int hs, ws, H, W;     // prediction size, ground-truth size
  float sy, sx;         // (float)hs / H, (float)ws / W
  float vscale;         // W / ws (1 when the sizes agree)
  float max_flow;
  int kind;             // 0: L1, 1: smooth L1 (beta = 1)
};

__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min((int)floorf(__fmul_rn((float)dst, scale)), in_size - 1);
}

// element loss of d = gt - p and its derivative with respect to p
__device__ __forceinline__ float loss_value(float d, int kind) {
  const float a = fabsf(d);
  return (kind == 1 && a < 1.f) ? 0.5f * d * d : (kind == 1 ? a - 0.5f : a);
}
__device__ __forceinline__ float loss_dpred(float d, int kind) {
  if (kind == 1 && fabsf(d) < 1.f) return -d;
  return d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
}

template <typename T>
__device__ __forceinline__ bool pixel_valid(const T* __restrict__ gt, const float* __restrict__ valid, int64_t p,
                                            float max_flow, float& g) {
  g = to_f(gt[p]);
  // flow_mag = sqrt(gt ** 2) for a one-channel flow (loss.py:55)
  return valid[p] >= 0.5f && sqrtf(__fmul_rn(g, g)) < max_flow;
}

// partial[b] = {sum of the element loss over valid pixels, valid count, non-finite predictions (all pixels),
//               infinite ground truth inside the mask}
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
seqloss_partial_kernel(const T* __restrict__ pred, const T* __restrict__ gt, const float* __restrict__ valid,
                       double* __restrict__ partial, int64_t total, LossGeom g) {
  // a block walks image rows (index arithmetic once per row), its threads the pixels of the row; counts stay
  // in 32-bit registers and the loss sum in fp32 per row (<= a few thousand terms) before joining the fp64 total
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t rows = total / max(g.W, 1);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int64_t n = row / g.H;
    const int y = (int)(row - n * g.H);
    const T* __restrict__ prow = pred + (n * g.hs + nearest_src(y, g.sy, g.hs)) * (int64_t)g.ws;
    float lsum = 0.f;
    int cnt = 0, bad = 0, infs = 0;
    for (int x = threadIdx.x; x < g.W; x += RED_THREADS) {
      const float raw = to_f(prow[nearest_src(x, g.sx, g.ws)]);
      float gv;
      const bool ok = pixel_valid(gt, valid, row * g.W + x, g.max_flow, gv);
      bad += !isfinite(raw);
      if (ok) {
        const float pv = to_f(from_f<T>(__fmul_rn(raw, g.vscale)));
        lsum += loss_value(__fsub_rn(gv, pv), g.kind);
        cnt += 1;
        infs += isinf(gv);
      }
    }
    acc[0] += (double)lsum; acc[1] += (double)cnt; acc[2] += (double)bad; acc[3] += (double)infs;
  }
  const bool no[4] = {false, false, false, false};
  block_reduce_store<4>(acc, no, no, partial + (int64_t)blockIdx.x * RED_SLOTS);
}

// result = {mean, count, non-finite predictions, infinite valid ground truth}
__global__ void __launch_bounds__(RED_THREADS)
seqloss_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ result) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += RED_THREADS)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += partial[(int64_t)b * RED_SLOTS + k];
  const bool no[4] = {false, false, false, false};
  __shared__ double out[4];
  block_reduce_store<4>(acc, no, no, out);
  __syncthreads();
  if (threadIdx.x == 0) {
    result[0] = out[0] / out[1];          // mean over an empty mask is NaN, as in torch
    result[1] = out[1]; result[2] = out[2]; result[3] = out[3];
  }
}

// gpred[n, sy, sx] = gmean / count * vscale * sum over the ground-truth pixels resized from (sy, sx) of dl/dp
template <typename T>
__global__ void __launch_bounds__(256)
seqloss_bwd_kernel(const float* __restrict__ gmean, const double* __restrict__ result, const T* __restrict__ pred,
                   const T* __restrict__ gt, const float* __restrict__ valid, T* __restrict__ gpred, LossGeom g) {
  const int64_t row = blockIdx.x;                  // prediction rows
  const int sy = (int)(row % g.hs);
  const int64_t n = row / g.hs;
  const int sx = blockIdx.y * 256 + threadIdx.x;
  if (sx >= g.ws) return;
  const float pv = to_f(from_f<T>(__fmul_rn(to_f(pred[row * g.ws + sx]), g.vscale)));
  const float iy = 1.f / g.sy, ix = 1.f / g.sx;
  int ylo = max((int)floorf((float)sy * iy) - 1, 0), yhi = min((int)ceilf((float)(sy + 1) * iy) + 1, g.H - 1);
  int xlo = max((int)floorf((float)sx * ix) - 1, 0), xhi = min((int)ceilf((float)(sx + 1) * ix) + 1, g.W - 1);
  if (sy == g.hs - 1) yhi = g.H - 1;               // clamped indices land on the last row / column
  if (sx == g.ws - 1) xhi = g.W - 1;
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    if (nearest_src(y, g.sy, g.hs) != sy) continue;
    for (int x = xlo; x <= xhi; ++x) {
      if (nearest_src(x, g.sx, g.ws) != sx) continue;
      const int64_t p = (n * g.H + y) * (int64_t)g.W + x;
      float gv;
      if (pixel_valid(gt, valid, p, g.max_flow, gv)) acc += loss_dpred(__fsub_rn(gv, pv), g.kind);
    }
  }
  const float scale = (float)((double)gmean[0] / result[1]) * g.vscale;
  gpred[row * g.ws + sx] = from_f<T>(acc * scale);
}

// partial = {sum epe, count, <0.5, <1, <3, <5, min pred[0], max pred[0]}
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
metrics_partial_kernel(const T* __restrict__ gt, const T* __restrict__ pred, const float* __restrict__ valid,
                       double* __restrict__ partial, int64_t total, int C, int64_t hw, int64_t wdim) {
  double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
  const int W = (int)wdim;
  const int64_t rows = total / max(W, 1), H = hw / max(W, 1);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int64_t n = row / H;
    const int64_t r0 = (row - n * H) * W;            // offset of the row inside its (H, W) plane
    float esum = 0.f, mn = INFINITY, mx = -INFINITY;
    int cnt = 0, c05 = 0, c1 = 0, c3 = 0, c5 = 0;
    for (int x = threadIdx.x; x < W; x += RED_THREADS) {
      float sq = 0.f;
      for (int c = 0; c < C; ++c) {
        const int64_t q = (n * C + c) * hw + r0 + x;
        const float pv = to_f(pred[q]);
        const float d = to_f(from_f<T>(__fsub_rn(pv, to_f(gt[q]))));
        sq = c == 0 ? to_f(from_f<T>(__fmul_rn(d, d))) : to_f(from_f<T>(__fadd_rn(sq, to_f(from_f<T>(__fmul_rn(d, d))))));
        if (n == 0) { mn = nan_minf(mn, pv); mx = nan_maxf(mx, pv); }
      }
      if (valid[row * W + x] >= 0.5f) {
        const float e = to_f(from_f<T>(sqrtf(sq)));
        esum += e; cnt += 1;
        c05 += e < 0.5f; c1 += e < 1.f; c3 += e < 3.f; c5 += e < 5.f;
      }
    }
    acc[0] += (double)esum; acc[1] += (double)cnt; acc[2] += (double)c05; acc[3] += (double)c1;
    acc[4] += (double)c3; acc[5] += (double)c5;
    acc[6] = nan_min(acc[6], (double)mn); acc[7] = nan_max(acc[7], (double)mx);
  }
  const bool mn[8] = {false, false, false, false, false, false, true, false};
  const bool mx[8] = {false, false, false, false, false, false, false, true};
  block_reduce_store<8>(acc, mn, mx, partial + (int64_t)blockIdx.x * RED_SLOTS);
}

// result = {epe, 0.5px, 1px, 3px, 5px, min, max, count}  (the order of the reference's dict)
__global__ void __launch_bounds__(RED_THREADS)
metrics_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ result) {
  double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
  for (int b = threadIdx.x; b < nblocks; b += RED_THREADS) {
    const double* s = partial + (int64_t)b * RED_SLOTS;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += s[k];
    acc[6] = nan_min(acc[6], s[6]); acc[7] = nan_max(acc[7], s[7]);
  }
  const bool mn[8] = {false, false, false, false, false, false, true, false};
  const bool mx[8] = {false, false, false, false, false, false, false, true};
  __shared__ double out[8];
  block_reduce_store<8>(acc, mn, mx, out);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double cnt = out[1];
    result[0] = out[0] / cnt;
    result[1] = out[2] / cnt; result[2] = out[3] / cnt; result[3] = out[4] / cnt; result[4] = out[5] / cnt;
    result[5] = out[6]; result[6] = out[7]; result[7] = cnt;
  }
}

static int reduce_blocks(int64_t rows) {
  return (int)std::min<int64_t>(RED_MAX_BLOCKS, std::max<int64_t>(1, rows));
}

static int loss_geom(int64_t N, int64_t hs, int64_t ws, int64_t H, int64_t W, float max_flow, int kind, LossGeom& g) {
  if (N < 0 || hs < 0 || ws < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  const int64_t lim = 1 << 24;
  if (hs > lim || ws > lim || H > lim || W > lim) return RSM_ERR_INVALID_SHAPE;
  if (N * H * W > 0 && hs * ws == 0) return RSM_ERR_INVALID_SHAPE;
  if (kind != 0 && kind != 1) return RSM_ERR_UNSUPPORTED_CONFIG;
  g.hs = (int)hs; g.ws = (int)ws; g.H = (int)H; g.W = (int)W;
  g.sy = H > 0 ? (float)hs / (float)H : 0.f;
  g.sx = W > 0 ? (float)ws / (float)W : 0.f;
  g.vscale = (hs == H && ws == W) || ws == 0 ? 1.f : (float)((double)W / (double)ws);
  g.max_flow = max_flow; g.kind = kind;
  return RSM_OK;
}

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

extern "C" int rsm_seqloss_fwd(const void* pred, const void* gt, const float* valid, double* workspace, double* result,
                               int64_t N, int64_t hs, int64_t ws, int64_t H, int64_t W, float max_flow, int kind,
                               int dtype, int device, void* stream) {
  LossGeom g;
  if (int rc = loss_geom(N, hs, ws, H, W, max_flow, kind, g)) return rc;
  if (!result || !workspace) return RSM_ERR_NULL_POINTER;
  const int64_t total = N * H * W;
  if (total > 0 && (!pred || !gt || !valid)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  const int nb = reduce_blocks(N * H);
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    seqloss_partial_kernel<T><<<nb, RED_THREADS, 0, st>>>((const T*)pred, (const T*)gt, valid, workspace, total, g);
    if (int rc = finish_launch("rsm_seqloss_fwd(partial)")) return rc;
    seqloss_final_kernel<<<1, RED_THREADS, 0, st>>>(workspace, nb, result);
    return finish_launch("rsm_seqloss_fwd(final)");
  });
}

extern "C" int rsm_seqloss_bwd(const float* gmean, const double* result, const void* pred, const void* gt,
                               const float* valid, void* gpred, int64_t N, int64_t hs, int64_t ws, int64_t H, int64_t W,
                               float max_flow, int kind, int dtype, int device, void* stream) {
  LossGeom g;
  if (int rc = loss_geom(N, hs, ws, H, W, max_flow, kind, g)) return rc;
  if (N * hs * ws == 0) return RSM_OK;
  if (!gmean || !result || !pred || !gpred || (N * H * W > 0 && (!gt || !valid))) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (N * hs > 2147483647LL) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const dim3 grid((unsigned)(N * hs), (unsigned)ceil_div(ws, 256));
    seqloss_bwd_kernel<T><<<grid, 256, 0, st>>>(gmean, result, (const T*)pred, (const T*)gt, valid, (T*)gpred, g);
    return finish_launch("rsm_seqloss_bwd");
  });
}

extern "C" int rsm_flow_metrics(const void* gt, const void* pred, const float* valid, double* workspace, double* result,
                                int64_t N, int64_t C, int64_t H, int64_t W, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return RSM_ERR_INVALID_SHAPE;
  if (!result || !workspace) return RSM_ERR_NULL_POINTER;
  const int64_t total = N * H * W;
  if (total * C > 0 && (!pred || !gt || !valid)) return RSM_ERR_NULL_POINTER;
  if (C > 65536) return RSM_ERR_INVALID_SHAPE;
  RSM_COMMON_CHECKS(dtype)
  const int nb = reduce_blocks(N * H);
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    metrics_partial_kernel<T><<<nb, RED_THREADS, 0, st>>>((const T*)gt, (const T*)pred, valid, workspace, C > 0 ? total : 0,
                                                          (int)C, H * W, W);
    if (int rc = finish_launch("rsm_flow_metrics(partial)")) return rc;
    metrics_final_kernel<<<1, RED_THREADS, 0, st>>>(workspace, nb, result);
    return finish_launch("rsm_flow_metrics(final)");
  });
}
