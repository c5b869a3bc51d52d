// Pre / post steps either side of the path, forward and adjoint.
//
// prepare:  model/mobile_stereo_net.py:121-130 (= _v2.py:194-203, _v3.py:296-305, mobile_disp_net_c.py:339-351,
//           _v4.py:433-434 without the pad): 2 * (img / 255) - 1, then F.pad(right / bottom, zeros) up to a multiple
//           of the model's alignment -- one pass, every op rounded like the reference's tensor ops (no FMA
//           contraction; 16-bit tensors round after each op as torch does).
// finalize: model/mobile_stereo_net.py:154-159 (= _v2.py:227-232; nearest) and mobile_disp_net_c.py:223-234 + :408-411
//           (bilinear, align_corners=False): F.interpolate(disp * scale, padded size)[:, :, :h, :w] * -1 -- the
//           four tensor ops (scale, resize, crop copy, negate) as one gather; the resized-but-cropped-away pixels
//           are never computed.  Source indices follow ATen's upsample kernels: scale = (float)in / out;
//           nearest: min((int)floorf(dst * scale), in - 1); linear: max(scale * (dst + 0.5) - 0.5, 0).
// Both are HBM / launch-bound element-wise work: algorithmic bytes = input read once + output written once.
// Adjoints are gathers per input element (deterministic, no atomics).
#include "rsm_common.cuh"

namespace rsm {

template <typename T> __device__ __forceinline__ float round_through(float v) { return to_f(from_f<T>(v)); }
template <> __device__ __forceinline__ float round_through<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ float normalise_pixel(float v) {
  // `img / 255.0` on a CUDA tensor is ATen's div-by-a-CPU-scalar kernel: a * (1 / b) with the reciprocal rounded to
  // fp32 once (BinaryDivTrueKernel.cu), not a true division -- the two differ by one ulp for about a third of the
  // 8-bit pixel values.  The drop-in follows the DEVICE convention, so that the patched model is bit-identical to the
  // unpatched model on the same GPU (a one-ulp change of the input moves a randomly initialised BatchNorm-in-train-mode
  // network's gradients by ~10 %: tools/diag_v4_train.py).  The CPU goldens (true division) agree within one ulp.
  const float a = round_through<T>(__fmul_rn(v, 1.0f / 255.0f));
  const float b = round_through<T>(__fmul_rn(2.f, a));
  return round_through<T>(__fsub_rn(b, 1.f));
}

// One thread per group of VEC output pixels, two groups per thread (independent loads in flight); the flat index
// is decoded with 32-bit divisions.  VEC = 4: 16-byte (fp32) / 8-byte (16-bit) stores, Wp % 4 == 0; the loads are
// vectors too when LDVEC (W % 4 == 0 and an aligned base), else four scalars (a raw 375 x 1242 KITTI frame).
template <typename T, int VEC, bool LDVEC>
__global__ void __launch_bounds__(256)
prepare_fwd_kernel(const T* __restrict__ img, T* __restrict__ out, int H, int W, int Hp, int Wp, int64_t groups) {
  const int rowg = Wp / VEC;
  const int64_t half = (groups + 1) >> 1;
  const int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i0 >= half) return;
  T v[2][VEC];
  bool ok[2][VEC];
  int64_t idx[2] = {i0, i0 + half};
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (idx[u] >= groups) break;
    const int64_t row = idx[u] / rowg;
    const int x = (int)(idx[u] - row * rowg) * VEC;
    const int64_t plane = row / Hp;
    const int y = (int)(row - plane * Hp);
    const T* s = img + (plane * H + y) * (int64_t)W + x;
    if constexpr (LDVEC) {
      const bool inside = y < H && x < W;          // W % 4 == 0: a vector is inside or outside as a whole
#pragma unroll
      for (int i = 0; i < VEC; ++i) ok[u][i] = inside;
      if (inside) {
        if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(v[u]) = *reinterpret_cast<const uint4*>(s);
        else *reinterpret_cast<uint2*>(v[u]) = *reinterpret_cast<const uint2*>(s);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        ok[u][i] = y < H && x + i < W;
        if (ok[u][i]) v[u][i] = s[i];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (idx[u] >= groups) break;
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[u][i] = from_f<T>(ok[u][i] ? normalise_pixel<T>(to_f(v[u][i])) : 0.f);
    T* o = out + idx[u] * VEC;
    if constexpr (VEC == 1) *o = v[u][0];
    else if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(v[u]);
    else *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(v[u]);
  }
}

// gimg = (gout[:, :, :H, :W] * 2) / 255, rounded like autograd's MulBackward / DivBackward chain
template <typename T>
__global__ void __launch_bounds__(128)
prepare_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gimg, int H, int W, int Hp, int Wp) {
  const int64_t row = blockIdx.x;                  // input rows
  const int y = (int)(row % H);
  const int64_t plane = row / H;
  const int x = blockIdx.y * 128 + threadIdx.x;
  if (x >= W) return;
  const float g = to_f(gout[(plane * Hp + y) * (int64_t)Wp + x]);
  gimg[row * W + x] = from_f<T>(round_through<T>(__fmul_rn(round_through<T>(__fmul_rn(g, 2.f)), 1.0f / 255.0f)));
}

enum : int { RESIZE_NEAREST = 0, RESIZE_BILINEAR = 1 };

struct Taps { int i0, i1; float l0, l1; };

template <int MODE>
__device__ __forceinline__ Taps resize_taps(int dst, float scale, int in_size) {
  Taps t;
  if constexpr (MODE == RESIZE_NEAREST) {
    t.i0 = t.i1 = min((int)floorf(__fmul_rn((float)dst, scale)), in_size - 1);
    t.l0 = 1.f; t.l1 = 0.f;
  } else {
    const float src = fmaxf(__fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f), 0.f);
    t.i0 = min((int)src, in_size - 1);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = fminf(fmaxf(src - (float)t.i0, 0.f), 1.f);
    t.l0 = 1.f - t.l1;
  }
  return t;
}

struct FinalGeom {
  int hs, ws;       // source (low-resolution disparity) size
  int Hp, Wp;       // size F.interpolate resizes to (the padded image)
  int h, w;         // crop = output size
  float sy, sx;     // (float)hs / Hp, (float)ws / Wp
  float vscale;     // value scale applied before the resize (Wp / ws in the reference)
};

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
finalize_fwd_kernel(const T* __restrict__ disp, T* __restrict__ out, FinalGeom g) {
  const int64_t row = blockIdx.x;                  // output rows
  const int y = (int)(row % g.h);
  const int64_t plane = row / g.h;
  const int x = blockIdx.y * 256 + threadIdx.x;
  if (x >= g.w) return;
  const T* __restrict__ src = disp + plane * (int64_t)g.hs * g.ws;
  const Taps ty = resize_taps<MODE>(y, g.sy, g.hs), tx = resize_taps<MODE>(x, g.sx, g.ws);
  auto val = [&](int yy, int xx) { return round_through<T>(__fmul_rn(to_f(src[(int64_t)yy * g.ws + xx]), g.vscale)); };
  float v;
  if constexpr (MODE == RESIZE_NEAREST) {
    v = val(ty.i0, tx.i0);
  } else {
    const float top = __fadd_rn(__fmul_rn(tx.l0, val(ty.i0, tx.i0)), __fmul_rn(tx.l1, val(ty.i0, tx.i1)));
    const float bot = __fadd_rn(__fmul_rn(tx.l0, val(ty.i1, tx.i0)), __fmul_rn(tx.l1, val(ty.i1, tx.i1)));
    v = round_through<T>(__fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot)));
  }
  out[row * g.w + x] = from_f<T>(-v);
}

// candidate destination range whose taps can touch source index s (then every candidate is tested exactly)
template <int MODE>
__device__ __forceinline__ void dst_range(int s, float scale, int out_size, int& lo, int& hi) {
  const float inv = 1.f / scale;
  if constexpr (MODE == RESIZE_NEAREST) {
    lo = (int)floorf((float)s * inv) - 1;
    hi = (int)ceilf((float)(s + 1) * inv) + 1;
  } else {
    lo = (int)floorf(((float)s - 0.5f) * inv - 0.5f) - 1;
    hi = (int)ceilf(((float)s + 1.5f) * inv - 0.5f) + 1;
  }
  lo = max(lo, 0);
  hi = min(hi, out_size - 1);
}

// gdisp[sy, sx] = -vscale * sum over the output pixels (inside the crop) whose taps hit (sy, sx)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
finalize_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gdisp, FinalGeom g) {
  const int64_t row = blockIdx.x;                  // source rows
  const int sy = (int)(row % g.hs);
  const int64_t plane = row / g.hs;
  const int sx = blockIdx.y * 256 + threadIdx.x;
  if (sx >= g.ws) return;
  const T* __restrict__ go = gout + plane * (int64_t)g.h * g.w;
  int ylo, yhi, xlo, xhi;
  dst_range<MODE>(sy, g.sy, min(g.h, g.Hp), ylo, yhi);
  dst_range<MODE>(sx, g.sx, min(g.w, g.Wp), xlo, xhi);
  // the last source row / column also collects every clamped index (nearest: min(., in - 1))
  if (sy == g.hs - 1) yhi = min(g.h, g.Hp) - 1;
  if (sx == g.ws - 1) xhi = min(g.w, g.Wp) - 1;
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    const Taps ty = resize_taps<MODE>(y, g.sy, g.hs);
    const float wy = (ty.i0 == sy ? ty.l0 : 0.f) + (ty.i1 == sy && ty.i1 != ty.i0 ? ty.l1 : 0.f) +
                     (MODE == RESIZE_BILINEAR && ty.i1 == ty.i0 && ty.i0 == sy ? ty.l1 : 0.f);
    if (wy == 0.f) continue;
    float racc = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      const Taps tx = resize_taps<MODE>(x, g.sx, g.ws);
      const float wx = (tx.i0 == sx ? tx.l0 : 0.f) + (tx.i1 == sx && tx.i1 != tx.i0 ? tx.l1 : 0.f) +
                       (MODE == RESIZE_BILINEAR && tx.i1 == tx.i0 && tx.i0 == sx ? tx.l1 : 0.f);
      if (wx != 0.f) racc = fmaf(wx, to_f(go[(int64_t)y * g.w + x]), racc);
    }
    acc = fmaf(wy, racc, acc);
  }
  gdisp[row * g.ws + sx] = from_f<T>(-g.vscale * acc);
}

static bool grid_ok(int64_t blocks) { return blocks >= 0 && blocks <= 2147483647LL; }

}  // namespace rsm

using namespace rsm;

#define RSM_COMMON_CHECKS(dtype)                                         \
  if (!valid_dtype(dtype)) return RSM_ERR_UNSUPPORTED_DTYPE;             \
  DeviceGuard guard(device);                                             \
  if (!guard.ok) { set_cuda_error(cudaGetLastError(), __func__); return RSM_ERR_CUDA; } \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

extern "C" int rsm_prepare_fwd(const void* img, void* out, int64_t planes, int64_t H, int64_t W, int64_t Hp,
                               int64_t Wp, int dtype, int device, void* stream) {
  if (planes < 0 || H < 0 || W < 0 || Hp < H || Wp < W || Hp > (1 << 24) || Wp > (1 << 24)) return RSM_ERR_INVALID_SHAPE;
  if (planes * Hp * Wp == 0) return RSM_OK;
  if (!out || (!img && H * W > 0)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const size_t vb = 4 * sizeof(T);
    const bool stvec = Wp % 4 == 0 && aligned_to(out, vb);
    const bool ldvec = stvec && W % 4 == 0 && aligned_to(img, vb);
    const int64_t groups = planes * Hp * (Wp / (stvec ? 4 : 1));
    const int64_t blocks = ceil_div((groups + 1) / 2, 256);
    if (!grid_ok(blocks)) return (int)RSM_ERR_INVALID_SHAPE;
    auto launch = [&](auto kern) {
      kern<<<(unsigned)blocks, 256, 0, st>>>((const T*)img, (T*)out, (int)H, (int)W, (int)Hp, (int)Wp, groups);
    };
    if (ldvec) launch(prepare_fwd_kernel<T, 4, true>);
    else if (stvec) launch(prepare_fwd_kernel<T, 4, false>);
    else launch(prepare_fwd_kernel<T, 1, false>);
    return finish_launch("rsm_prepare_fwd");
  });
}

extern "C" int rsm_prepare_bwd(const void* gout, void* gimg, int64_t planes, int64_t H, int64_t W, int64_t Hp,
                               int64_t Wp, int dtype, int device, void* stream) {
  if (planes < 0 || H < 0 || W < 0 || Hp < H || Wp < W || Hp > (1 << 24) || Wp > (1 << 24)) return RSM_ERR_INVALID_SHAPE;
  if (planes * H * W == 0) return RSM_OK;
  if (!gout || !gimg) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(planes * H)) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const dim3 grid((unsigned)(planes * H), (unsigned)ceil_div(W, 128));
    prepare_bwd_kernel<T><<<grid, 128, 0, st>>>((const T*)gout, (T*)gimg, (int)H, (int)W, (int)Hp, (int)Wp);
    return finish_launch("rsm_prepare_bwd");
  });
}

static int finalize_geom(int64_t planes, int64_t hs, int64_t ws, int64_t Hp, int64_t Wp, int64_t h, int64_t w,
                         float vscale, int mode, FinalGeom& g) {
  if (planes < 0 || hs < 0 || ws < 0 || Hp < 0 || Wp < 0 || h < 0 || w < 0 || h > Hp || w > Wp) return RSM_ERR_INVALID_SHAPE;
  if (mode != RESIZE_NEAREST && mode != RESIZE_BILINEAR) return RSM_ERR_UNSUPPORTED_CONFIG;
  const int64_t lim = 1 << 24;                     // pixel indices must be exact in fp32
  if (hs > lim || ws > lim || Hp > lim || Wp > lim) return RSM_ERR_INVALID_SHAPE;
  if (planes * h * w > 0 && hs * ws == 0) return RSM_ERR_INVALID_SHAPE;   // nothing to resize from
  g.hs = (int)hs; g.ws = (int)ws; g.Hp = (int)Hp; g.Wp = (int)Wp; g.h = (int)h; g.w = (int)w;
  g.sy = Hp > 0 ? (float)hs / (float)Hp : 0.f;
  g.sx = Wp > 0 ? (float)ws / (float)Wp : 0.f;
  g.vscale = vscale;
  return RSM_OK;
}

extern "C" int rsm_finalize_fwd(const void* disp, void* out, int64_t planes, int64_t hs, int64_t ws, int64_t Hp,
                                int64_t Wp, int64_t h, int64_t w, float vscale, int mode, int dtype, int device,
                                void* stream) {
  FinalGeom g;
  if (int rc = finalize_geom(planes, hs, ws, Hp, Wp, h, w, vscale, mode, g)) return rc;
  if (planes * h * w == 0) return RSM_OK;
  if (!disp || !out) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(planes * h)) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const dim3 grid((unsigned)(planes * h), (unsigned)ceil_div(w, 256));
    if (mode == RESIZE_NEAREST) finalize_fwd_kernel<T, RESIZE_NEAREST><<<grid, 256, 0, st>>>((const T*)disp, (T*)out, g);
    else finalize_fwd_kernel<T, RESIZE_BILINEAR><<<grid, 256, 0, st>>>((const T*)disp, (T*)out, g);
    return finish_launch("rsm_finalize_fwd");
  });
}

extern "C" int rsm_finalize_bwd(const void* gout, void* gdisp, int64_t planes, int64_t hs, int64_t ws, int64_t Hp,
                                int64_t Wp, int64_t h, int64_t w, float vscale, int mode, int dtype, int device,
                                void* stream) {
  FinalGeom g;
  if (int rc = finalize_geom(planes, hs, ws, Hp, Wp, h, w, vscale, mode, g)) return rc;
  if (planes * hs * ws == 0) return RSM_OK;
  if (!gdisp || (!gout && h * w > 0)) return RSM_ERR_NULL_POINTER;
  RSM_COMMON_CHECKS(dtype)
  if (!grid_ok(planes * hs)) return RSM_ERR_INVALID_SHAPE;
  return RSM_DISPATCH_DTYPE(dtype, T, [&]() -> int {
    const dim3 grid((unsigned)(planes * hs), (unsigned)ceil_div(ws, 256));
    if (mode == RESIZE_NEAREST) finalize_bwd_kernel<T, RESIZE_NEAREST><<<grid, 256, 0, st>>>((const T*)gout, (T*)gdisp, g);
    else finalize_bwd_kernel<T, RESIZE_BILINEAR><<<grid, 256, 0, st>>>((const T*)gout, (T*)gdisp, g);
    return finish_launch("rsm_finalize_bwd");
  });
}
