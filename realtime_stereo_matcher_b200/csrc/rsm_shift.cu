// Shifted interweave stack: all D iterations of MobileStereoNetV4's per-disparity volume loop
// (model/mobile_stereo_net_v4.py:444-458: interweave_tensors(featL[..., i:], featR[..., :-i])) built in one
// pass, full width, zero where x < d.  Pure data movement, HBM-write bound (D * 4NCHW*e bytes out):
// one CTA per output row (d, n, ch, y); the source row goes to shared memory once and the W outputs are
// written with 128-bit streaming stores.  The adjoint is an atomic-free gather over d.
#include "rsm_common.cuh"

namespace rsm {

constexpr int kShiftThreads = 128;

template <typename T, int VEC>
__global__ void __launch_bounds__(kShiftThreads)
shift_interweave_fwd_kernel(FeatView L, FeatView R, T* __restrict__ out, int N, int C, int H, int W, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* srow = reinterpret_cast<T*>(smem_raw);
  // one CTA handles the (n, ch, y) row for ALL disparities: the source row is read once, D rows are written
  int64_t row = blockIdx.x;
  const int y = (int)(row % H); row /= H;
  const int ch = (int)(row % (2 * C));
  const int64_t n = row / (2 * C);
  const bool right = ch & 1;
  const FeatView& F = right ? R : L;
  const T* __restrict__ src = reinterpret_cast<const T*>(F.data) + n * F.sn + (int64_t)(ch >> 1) * F.sc + (int64_t)y * F.sh;
  for (int x = threadIdx.x; x < W; x += kShiftThreads) srow[x] = __ldg(src + (int64_t)x * F.sw);
  __syncthreads();
  const T zero = from_f<T>(0.f);
  const int WV = W / VEC;
  const int64_t dstride = (int64_t)N * 2 * C * H * W;
  T* __restrict__ o0 = out + ((n * 2 * C + ch) * H + y) * (int64_t)W;
  for (int d = 0; d < D; ++d) {
    T* __restrict__ o = o0 + d * dstride;
    for (int xv = threadIdx.x; xv < WV; xv += kShiftThreads) {
      const int x0 = xv * VEC;
      T vals[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int x = x0 + j;
        vals[j] = (x >= d) ? srow[right ? x - d : x] : zero;
      }
      if constexpr (VEC * sizeof(T) == 16) {
        Vec16<T> v;
#pragma unroll
        for (int j = 0; j < VEC; ++j) v.v[j] = vals[j];
        stcs16(o + x0, v);
      } else {
        o[x0] = vals[0];
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
shift_interweave_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gl, T* __restrict__ gr, int64_t total,
                            int N, int C, int H, int W, int D) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int c = (int)((i / ((int64_t)W * H)) % C);
  const int64_t n = i / ((int64_t)W * H * C);
  const int64_t dstride = (int64_t)N * 2 * C * H * W;
  const T* pl = gout + ((n * 2 * C + 2 * c) * H + y) * (int64_t)W + x;
  const T* pr = pl + (int64_t)H * W;
  float sl = 0.f, sr = 0.f;
  const int dl = min(x, D - 1);
  for (int d = 0; d <= dl; ++d) sl += to_f(__ldg(pl + d * dstride));
  const int dr = min(D - 1, W - 1 - x);
  for (int d = 0; d <= dr; ++d) sr += to_f(__ldg(pr + d * dstride + d));
  gl[i] = from_f<T>(sl);
  gr[i] = from_f<T>(sr);
}

}  // namespace rsm

using namespace rsm;

extern "C" int rsm_shift_interweave_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                                        int64_t W, int64_t D, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  if (N * C * H * W * D == 0) return RSM_OK;
  if (!left.data || !right.data || !out) return RSM_ERR_NULL_POINTER;
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (N * 2 * C * H > 2147483647LL || W > (1 << 24) || D > (1 << 24)) return RSM_ERR_INVALID_SHAPE;
  DeviceGuard guard(device);
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    constexpr int VEC = 16 / sizeof(T);
    const size_t smem = (size_t)W * sizeof(T);
    if (smem > 200 * 1024) return (int)RSM_ERR_UNSUPPORTED_CONFIG;
    const unsigned grid = (unsigned)(N * 2 * C * H);
    const bool vec = W % VEC == 0 && aligned_to(out, 16);
    if (vec) {
      auto k = shift_interweave_fwd_kernel<T, VEC>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, kShiftThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)N, (int)C, (int)H, (int)W, (int)D);
    } else {
      auto k = shift_interweave_fwd_kernel<T, 1>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, kShiftThreads, smem, st>>>(view_of(left), view_of(right), (T*)out, (int)N, (int)C, (int)H, (int)W, (int)D);
    }
    return finish_launch("rsm_shift_interweave_fwd");
  });
}

extern "C" int rsm_shift_interweave_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                                        int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || D < 0) return RSM_ERR_INVALID_SHAPE;
  const int64_t total = N * C * H * W;
  if (total == 0) return RSM_OK;
  if (!gleft || !gright || (D > 0 && !gout)) return RSM_ERR_NULL_POINTER;
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;
  if (ceil_div(total, 256) > 2147483647LL) return RSM_ERR_INVALID_SHAPE;
  DeviceGuard guard(device);
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    shift_interweave_bwd_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const T*)gout, (T*)gleft, (T*)gright, total, (int)N, (int)C, (int)H, (int)W, (int)D);
    return finish_launch("rsm_shift_interweave_bwd");
  });
}
