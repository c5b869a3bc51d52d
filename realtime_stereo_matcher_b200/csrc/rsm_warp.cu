// Refinement warp: warp_by_flow_map, model/mobile_stereo_net_v2.py:59-96
// (= model/mobile_stereo_net_v3.py:60-97, tools/warp.py:5-42), forward and adjoint.
//
// The reference builds grid = pixel index - flow, normalises it with (size - 1) and samples with
// F.grid_sample(bilinear, zeros, align_corners=False), whose un-normalisation uses size, so the source
// coordinate is ix = (x - f0) * W / (W - 1) - 0.5 and -- even for a one-channel flow -- iy = y * H / (H - 1)
// - 0.5 is not an integer row: a genuine 2-D bilinear gather.  The coordinates are evaluated with the
// reference's own fp32 operation sequence (no FMA contraction), once per pixel; the channel loop then moves
// 4 taps per channel.  HBM/L2-bound: algorithmic bytes = image + flow read once, output written once.
// Adjoint: gimage by fp32 atomics (like ATen's grid_sampler backward), gflow summed per pixel in registers.
#include "rsm_common.cuh"

namespace rsm {

struct WarpTaps {
  int x0, y0;          // north-west tap
  float wx1, wy1;      // weight of the east / south taps (west / north = 1 - w)
  bool inx0, inx1, iny0, iny1;
};

__device__ __forceinline__ float warp_unnorm(float idx_minus_flow, int size) {
  // 2 * g / (size - 1) - 1, then ((. + 1) * size - 1) / 2, each rounded like the reference's tensor ops
  const float gn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, idx_minus_flow), (float)size - 1.f), 1.f);
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gn, 1.f), (float)size), 1.f), 2.f);
}

template <typename T>
__device__ __forceinline__ WarpTaps warp_taps(const T* __restrict__ flow_n, int cf, int y, int x, int H, int W) {
  const int64_t hw = (int64_t)H * W, p = (int64_t)y * W + x;
  const float ix = warp_unnorm(__fsub_rn((float)x, to_f(flow_n[p])), W);
  const float iy = warp_unnorm(cf == 2 ? __fsub_rn((float)y, to_f(flow_n[hw + p])) : (float)y, H);
  const float fx = floorf(ix), fy = floorf(iy);
  WarpTaps t;
  // out-of-range (or non-finite) coordinates: every tap fails its bounds test; keep the ints harmless
  t.x0 = (fx >= -2.f && fx <= (float)W) ? (int)fx : -2;
  t.y0 = (fy >= -2.f && fy <= (float)H) ? (int)fy : -2;
  t.wx1 = ix - fx; t.wy1 = iy - fy;
  t.inx0 = t.x0 >= 0 && t.x0 < W; t.inx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.iny0 = t.y0 >= 0 && t.y0 < H; t.iny1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  return t;
}

constexpr int WARP_CCH = 8;   // channels per thread: blockIdx.y walks the channel chunks (more loads in flight
                              // than one thread per pixel; the coordinates are cheap to recompute)
template <typename T>
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const T* __restrict__ image, const T* __restrict__ flow, T* __restrict__ out, int64_t total, int C,
                int H, int W, int cf) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;      // (n, y, x)
  if (i >= total) return;
  const int cbeg = blockIdx.y * WARP_CCH, cend = min(C, cbeg + WARP_CCH);
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int64_t n = i / ((int64_t)W * H), hw = (int64_t)H * W;
  const WarpTaps t = warp_taps(flow + n * cf * hw, cf, y, x, H, W);
  const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
  const float wnw = (t.inx0 && t.iny0) ? wx0 * wy0 : 0.f, wne = (t.inx1 && t.iny0) ? t.wx1 * wy0 : 0.f;
  const float wsw = (t.inx0 && t.iny1) ? wx0 * t.wy1 : 0.f, wse = (t.inx1 && t.iny1) ? t.wx1 * t.wy1 : 0.f;
  // clamped tap addresses (weights of outside taps are zero)
  const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
  const int ya = min(max(t.y0, 0), H - 1), yb = min(max(t.y0 + 1, 0), H - 1);
  const int64_t onw = (int64_t)ya * W + xa, one = (int64_t)ya * W + xb, osw = (int64_t)yb * W + xa, ose = (int64_t)yb * W + xb;
  const T* __restrict__ src = image + (n * C + cbeg) * hw;
  T* __restrict__ dst = out + (n * C + cbeg) * hw + (int64_t)y * W + x;
#pragma unroll 8
  for (int c = cbeg; c < cend; ++c, src += hw, dst += hw) {
    const float v = wnw * to_f(__ldg(src + onw)) + wne * to_f(__ldg(src + one)) + wsw * to_f(__ldg(src + osw)) +
                    wse * to_f(__ldg(src + ose));
    *dst = from_f<T>(v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ image, const T* __restrict__ flow,
                float* __restrict__ gimage, T* __restrict__ gflow, int64_t total, int C, int H, int W, int cf) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int64_t n = i / ((int64_t)W * H), hw = (int64_t)H * W;
  const WarpTaps t = warp_taps(flow + n * cf * hw, cf, y, x, H, W);
  const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
  const bool bnw = t.inx0 && t.iny0, bne = t.inx1 && t.iny0, bsw = t.inx0 && t.iny1, bse = t.inx1 && t.iny1;
  const int64_t onw = (int64_t)t.y0 * W + t.x0, one = onw + 1, osw = onw + W, ose = onw + W + 1;
  const T* __restrict__ src = image + n * C * hw;
  const T* __restrict__ g = gout + n * C * hw + (int64_t)y * W + x;
  float* __restrict__ gi = gimage + n * C * hw;
  float gix = 0.f, giy = 0.f;
  for (int c = 0; c < C; ++c, src += hw, g += hw, gi += hw) {
    const float go = to_f(*g);
    const float vnw = bnw ? to_f(__ldg(src + onw)) : 0.f, vne = bne ? to_f(__ldg(src + one)) : 0.f;
    const float vsw = bsw ? to_f(__ldg(src + osw)) : 0.f, vse = bse ? to_f(__ldg(src + ose)) : 0.f;
    if (gimage) {                       // NULL: the image needs no gradient (v2 warps the raw right image)
      if (bnw) atomicAdd(gi + onw, go * wx0 * wy0);
      if (bne) atomicAdd(gi + one, go * t.wx1 * wy0);
      if (bsw) atomicAdd(gi + osw, go * wx0 * t.wy1);
      if (bse) atomicAdd(gi + ose, go * t.wx1 * t.wy1);
    }
    gix = fmaf(go, (vne - vnw) * wy0 + (vse - vsw) * t.wy1, gix);
    giy = fmaf(go, (vsw - vnw) * wx0 + (vse - vne) * t.wx1, giy);
  }
  if (gflow) {
    T* gf = gflow + n * cf * hw + (int64_t)y * W + x;
    gf[0] = from_f<T>(-gix * ((float)W / ((float)W - 1.f)));
    if (cf == 2) gf[hw] = from_f<T>(-giy * ((float)H / ((float)H - 1.f)));
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

extern "C" int rsm_warp_fwd(const void* image, const void* flow, void* out, int64_t N, int64_t C, int64_t H,
                            int64_t W, int flow_channels, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || (flow_channels != 1 && flow_channels != 2)) return RSM_ERR_INVALID_SHAPE;
  if (H > (1 << 24) || W > (1 << 24)) return RSM_ERR_INVALID_SHAPE;     // pixel indices must be exact in fp32
  const int64_t total = N * H * W;
  if (total * C == 0) return RSM_OK;
  if (!image || !flow || !out) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(ceil_div(total, 256)) || ceil_div(C, WARP_CCH) > 65535) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const dim3 grid((unsigned)ceil_div(total, 256), (unsigned)ceil_div(C, WARP_CCH));
    warp_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)image, (const T*)flow, (T*)out, total, (int)C, (int)H, (int)W,
                                             flow_channels);
    return finish_launch("rsm_warp_fwd");
  });
}

extern "C" int rsm_warp_bwd(const void* gout, const void* image, const void* flow, float* gimage, void* gflow,
                            int64_t N, int64_t C, int64_t H, int64_t W, int flow_channels, int dtype, int device,
                            void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || (flow_channels != 1 && flow_channels != 2)) return RSM_ERR_INVALID_SHAPE;
  if (H > (1 << 24) || W > (1 << 24)) return RSM_ERR_INVALID_SHAPE;
  const int64_t total = N * H * W;
  if (total == 0) return RSM_OK;
  if ((!gimage && !gflow) || (C > 0 && (!gout || !image)) || !flow) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(ceil_div(total, 256))) return RSM_ERR_INVALID_SHAPE;
  if (gimage && C > 0 && cudaMemsetAsync(gimage, 0, (size_t)(total * C) * sizeof(float), st) != cudaSuccess)
    return finish_launch("rsm_warp_bwd(memset)");
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    warp_bwd_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const T*)gout, (const T*)image, (const T*)flow,
                                                                      gimage, (T*)gflow, total, (int)C, (int)H, (int)W,
                                                                      flow_channels);
    return finish_launch("rsm_warp_bwd");
  });
}
