// Adjoint of the group-wise volume (N,G,H,W,D), D innermost (cost_volume/groupwise.py:39-55 under autograd):
//
//   gL[c,x]  = s * sum_d gV[g(c), x, d]      * R[c, x - d]
//   gR[c,x'] = s * sum_d gV[g(c), x' + d, d] * L[c, x' + d]            s = 1 / (channels per group)
//
// Work unit = one (n, group, y) ROW of the gradient: W x D contiguous values.  A persistent CTA walks a contiguous
// run of rows and streams them through a ring of shared-memory slabs (16 disparities of fp32, 32 of a 16-bit
// gradient: 64 contiguous bytes per pixel, the granule HBM delivers -- 32-byte pieces measured 1.5x the algorithmic
// DRAM traffic) with plain 16-byte cp.async: no registers, no transposition, the next slab (of this row or the
// next one) in flight under the FMAs of the current one; the next row's features arrive the same way into the
// second feature buffer.  A one-row-per-CTA, register-staged form of this kernel was bound by the latency of its
// fetches (2.2 TB/s) and spent 42 % of its instructions on per-CTA set-up.
//
// Slab layout: sG[pixel quad][pixel in quad (4)][d] with a quad pitch of 4 (2 for 16-bit) words more than the data:
//   * a pixel's disparities stay contiguous (the copy is a copy), and the 4 x 4 block gV[x..x+3][d..d+3] a
//     register tile needs is four LDS.128, one per pixel -- the transposition happens in the register names;
//   * every such load is at a COMPILE-TIME offset from one per-thread base;
//   * pitch 68 = 4 (mod 32): the LDS.128 of eight neighbouring quads hit eight different bank quads
//     (16-bit: pitch 66 = 2 (mod 32) for the LDS.64 of sixteen neighbouring quads).
//
// A thread owns 4 pixels x all CPG channels of ONE side for the whole row (first half of the CTA: gL, second
// half: gR; halves are whole warps) and keeps that 4 x CPG tile in registers across the slabs: no split of the
// disparity range, no partial sums, nothing but the two operands is read from shared memory in the loop (left:
// 4 + 2 CPG aligned LDS.128 per 16 CPG FMAs; right: the gradient pixels x' + d .. x' + d + 3 of four consecutive d
// span 7 pixels, so 7 + 2 CPG).  The gradient is read from HBM exactly once and there is no halo: the right side
// needs gV[x' + d] only for x' + d < W, which is the same slab.  Both feature rows of the group sit in shared
// memory as fp32 with zero margins standing in for the x < d fill.  Atomic-free and deterministic (ascending d).
#pragma once

namespace rsm {

constexpr int GS_MAXT = 512;           // threads per CTA (2 sides x <= 256 pixel quads: W <= 1024)
constexpr int GS_UNROLL_CPG = 8;        // whole slabs are unrolled up to this many channels per group (16: register spills);
                                        // 16-bit gradients (8 quads per slab) up to 4: with 8 the unrolled form measured 10-17 % slower
constexpr int GS_STAGES = 2;           // slab stages of the ring: with three, only two CTAs fit an SM at W = 312 (measured 163 vs 134 us)

template <typename Tout> struct GsSlab {
  static constexpr bool F32 = sizeof(Tout) == 4;
  static constexpr int DC = F32 ? 16 : 32;        // disparities per slab (64 bytes per pixel)
  static constexpr int QP = F32 ? 68 : 66;        // words per pixel quad
  static constexpr int PB = F32 ? 16 : 8;         // bytes per cp.async piece (the 16-bit pitch is only 8-byte aligned)
  static constexpr int PPP = 64 / PB;             // pieces per pixel
};

struct GroupSlabGeom {
  int NQ;        // pixel quads per slab: ceil(W / 4) + 2 (quads past the row stay zero)
  int FPL;       // pitch of a left-feature row:  ceil4(W) + 8   (L[c][x] at x)
  int FPR;       // pitch of a right-feature row: DPAD + ceil4(W) (R[c][x] at DPAD + x)
  int DPAD;      // zero margin in front of the right features: DC * nslab + 4
  int half;      // threads per side: ceil32(ceil(W / 4))
  int nslab;     // ceil(D / DC)
  int fvec;      // feature rows can be copied 16 bytes at a time (unit stride, aligned rows, W % (16 / sizeof) == 0)
  int nfb;       // feature buffers: 2 (the next row's features arrive under this row's FMAs), or 1 where two do not fit
  int64_t rows;  // N * G * H
};

// A thread's share of the cp.async copy of a slab.  Piece u = threadIdx.x + k * blockDim.x -> pixel u / PPP, sub-piece
// u % PPP; blockDim is a multiple of 64, so the sub-piece is per-thread constant and pixel, source and destination
// advance by uniform steps in k: five instructions per piece, no divisions.
template <typename Tout>
struct GsCopy {
  using S = GsSlab<Tout>;
  static constexpr int EPP = S::PB / (int)sizeof(Tout);   // elements per piece (4)
  int64_t soff, sstep;   // element offset of piece k = 0 inside a row (at d0 = 0), elements between pieces k and k + 1
  uint32_t doff, dstep;  // byte offset of piece k = 0 inside a stage, bytes between pieces
  int p0, pstep, dsub;   // pixel of piece k = 0, pixels between pieces, first disparity of the sub-piece
  __device__ __forceinline__ GsCopy(int D) {
    const int t = threadIdx.x, bd = blockDim.x;
    p0 = t / S::PPP; pstep = bd / S::PPP; dsub = (t % S::PPP) * EPP;
    soff = (int64_t)p0 * D + dsub; sstep = (int64_t)pstep * D;
    doff = 4u * ((p0 >> 2) * S::QP + (p0 & 3) * 16) + (uint32_t)S::PB * (t % S::PPP);
    dstep = 4u * (pstep >> 2) * S::QP;
  }
  // disparities d0 .. d0 + DC of the row at `grow` -> the stage at shared address `sbase`.  Pixels >= W are never
  // written (the stages are zeroed once); a sub-piece past D is zeroed (only in the last slab when D % DC != 0).
  __device__ __forceinline__ void issue(uint32_t sbase, const Tout* __restrict__ grow, int W, int D, int d0) const {
    const Tout* src = grow + soff + d0;
    uint32_t dst = sbase + doff;
    if (d0 + dsub < D) {
#pragma unroll 4
      for (int p = p0; p < W; p += pstep, src += sstep, dst += dstep) {
        if constexpr (S::F32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      }
    } else {
      for (int p = p0; p < W; p += pstep, dst += dstep) {
        if constexpr (S::F32) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
        else asm volatile("st.shared.v2.b32 [%0], {%1, %1};" ::"r"(dst), "r"(0) : "memory");
      }
    }
  }
};

// the 4 disparities d0 + 4q .. + 3 of pixel i of the quad at sGq, as fp32
template <typename Tout>
__device__ __forceinline__ float4 gs_ld4(const float* __restrict__ sGq, int i, int q) {
  if constexpr (sizeof(Tout) == 4) {
    return *reinterpret_cast<const float4*>(sGq + 16 * i + 4 * q);
  } else {
    const uint2 t = *reinterpret_cast<const uint2*>(sGq + 16 * i + 2 * q);
    const float2 a = unpack2<Tout>(t.x), b = unpack2<Tout>(t.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

// one side's FMAs for disparity quads [Q0, Q1) of a slab; sGq = the slab at this thread's pixel quad
template <typename Tout, int CPG, int NQ>
__device__ __forceinline__ void gs_left(float (&acc)[CPG][4], const float* __restrict__ sGq, const float* __restrict__ wrow,
                                        int FPR, int Q0, int Q1) {
  // R[c][x + i - d], d = d0 + 4q + r: window w[0..8) = wrow[c * FPR - 4q ..], element 4 + i - r
  // (NQ > 0: the first NQ quads of the slab, unrolled -- every address is then an immediate offset)
#pragma unroll(NQ > 0 ? NQ : 1)
  for (int q = NQ > 0 ? 0 : Q0; q < (NQ > 0 ? NQ : Q1); ++q) {
    float gq[4][4];                                    // [r][i]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = gs_ld4<Tout>(sGq, i, q);
      gq[0][i] = t.x; gq[1][i] = t.y; gq[2][i] = t.z; gq[3][i] = t.w;
    }
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(wrow + j * FPR - 4 * q);
      const float4 w0 = wp[0], w1 = wp[1];
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(gq[r][i], w[4 + i - r], acc[j][i]);
    }
  }
}

// sGu / lrow = the slab / the left features at pixel x' + d0 (u = x' + d)
template <typename Tout, int CPG, int NQ>
__device__ __forceinline__ void gs_right(float (&acc)[CPG][4], const float* __restrict__ sGu, const float* __restrict__ lrow,
                                         int FPL, int ub, int W, int Q0, int Q1) {
  // u0 = x' + d0 + 4q: gV[d0 + 4q + r][u0 + r + i] is component r of pixel m = r + i of the two quads at u0;
  // L[c][u0 + r + i] is element r + i of the aligned octet at u0
  constexpr int QP = GsSlab<Tout>::QP;
#pragma unroll(NQ > 0 ? NQ : 1)
  for (int q = NQ > 0 ? 0 : Q0; q < (NQ > 0 ? NQ : Q1); ++q) {
    if (ub + 4 * q >= W) break;                        // nothing but zeros further right
    float p[4][4];                                     // [r][i]
#pragma unroll
    for (int m = 0; m < 7; ++m) {
      const float4 t = gs_ld4<Tout>(sGu + (q + (m >> 2)) * QP, m & 3, q);
      const float c[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (m - r >= 0 && m - r < 4) p[r][m - r] = c[r];
    }
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float4* lp = reinterpret_cast<const float4*>(lrow + j * FPL + 4 * q);
      const float4 l0 = lp[0], l1 = lp[1];
      const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(p[r][i], l[r + i], acc[j][i]);
    }
  }
}

// The group's 2 x CPG feature rows of image row y -> fp32 at sF ([CPG][FPL] left, then [CPG][FPR] right at DPAD + x).
// Only x < W is written: the zero margins are laid down once per CTA.  fp32 rows arrive by cp.async (the caller
// waits), 16-bit rows are widened on the way.
template <typename Tin, int CPG>
__device__ __forceinline__ void gs_stage_features(float* __restrict__ sF, const FeatView& L, const FeatView& R, int64_t n,
                                                  int c0, int y, int W, const GroupSlabGeom& sg) {
  constexpr int EPV = 16 / (int)sizeof(Tin);
  const Tin* __restrict__ pl = reinterpret_cast<const Tin*>(L.data) + n * L.sn + (int64_t)c0 * L.sc + (int64_t)y * L.sh;
  const Tin* __restrict__ pr = reinterpret_cast<const Tin*>(R.data) + n * R.sn + (int64_t)c0 * R.sc + (int64_t)y * R.sh;
  float* dl = sF;
  float* dr = sF + CPG * sg.FPL + sg.DPAD;
  if (sg.fvec) {
    const int x = EPV * threadIdx.x;                   // blockDim >= W / 2 >= W / EPV
    if (x >= W) return;
    pl += x; pr += x; dl += x; dr += x;
#pragma unroll
    for (int c = 0; c < CPG; ++c, pl += L.sc, pr += R.sc, dl += sg.FPL, dr += sg.FPR) {
      if constexpr (sizeof(Tin) == 4) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dl)), "l"(pl) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dr)), "l"(pr) : "memory");
      } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(pl)), b = __ldg(reinterpret_cast<const uint4*>(pr));
        const float2 a0 = unpack2<Tin>(a.x), a1 = unpack2<Tin>(a.y), a2 = unpack2<Tin>(a.z), a3 = unpack2<Tin>(a.w);
        const float2 b0 = unpack2<Tin>(b.x), b1 = unpack2<Tin>(b.y), b2 = unpack2<Tin>(b.z), b3 = unpack2<Tin>(b.w);
        *reinterpret_cast<float4*>(dl) = make_float4(a0.x, a0.y, a1.x, a1.y);
        *reinterpret_cast<float4*>(dl + 4) = make_float4(a2.x, a2.y, a3.x, a3.y);
        *reinterpret_cast<float4*>(dr) = make_float4(b0.x, b0.y, b1.x, b1.y);
        *reinterpret_cast<float4*>(dr + 4) = make_float4(b2.x, b2.y, b3.x, b3.y);
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < CPG; ++c, pl += L.sc, pr += R.sc, dl += sg.FPL, dr += sg.FPR) {
      for (int x = threadIdx.x; x < W; x += blockDim.x) {
        if constexpr (sizeof(Tin) == 4) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dl + x)),
                       "l"(pl + (int64_t)x * L.sw) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dr + x)),
                       "l"(pr + (int64_t)x * R.sw) : "memory");
        } else {
          dl[x] = to_f(__ldg(pl + (int64_t)x * L.sw));
          dr[x] = to_f(__ldg(pr + (int64_t)x * R.sw));
        }
      }
    }
  }
}

template <typename Tin, typename Tout, int CPG, int MAXT, int GS_NS>
__global__ void __launch_bounds__(MAXT) __maxnreg__(MAXT <= 192 ? 96 : 128)
groupwise_bwd_slab_kernel(const Tout* __restrict__ gout, FeatView L, FeatView R, Tin* __restrict__ gl,
                          Tin* __restrict__ gr, CorrGeom g, GroupSlabGeom sg) {
  using S = GsSlab<Tout>;
  extern __shared__ __align__(16) float smem[];
  const int slab = sg.NQ * S::QP + (sg.NQ * S::QP & 1);   // words per stage (16-byte multiple)
  const int fsz = CPG * (sg.FPL + sg.FPR);
  float* sG = smem;                                    // [GS_NS][NQ][4][DC] (quad pitch QP)
  float* sF = sG + GS_NS * (slab + (slab & 2));        // [nfb][ CPG x FPL | CPG x FPR ]
  const int stage = slab + (slab & 2);
  // this CTA's run of rows
  const int64_t r0 = sg.rows * blockIdx.x / gridDim.x, r1 = sg.rows * (blockIdx.x + 1) / gridDim.x;
  if (r0 >= r1) return;
  const int64_t WD = (int64_t)g.W * g.D;
  const GsCopy<Tout> cp(g.D);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sG);

  // ---- zero the stages (quads past the row are never written) and the feature buffers (margins)
  for (int i = threadIdx.x; i < (GS_NS * stage + sg.nfb * fsz) / 4; i += blockDim.x)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  // ---- prologue: features of the first row and the first GS_NS - 1 slabs, one cp.async group per slab
  const int nunits = (int)(r1 - r0) * sg.nslab;        // slabs of this CTA
  int iu = 0, is = 0;                                  // next slab to issue: unit index, slab of its row
  const Tout* __restrict__ irow = gout + r0 * WD;      // its row
  {
    const int y = (int)(r0 % g.H);
    const int64_t t = r0 / g.H;
    gs_stage_features<Tin, CPG>(sF, L, R, t / g.G, (int)(t % g.G) * CPG, y, g.W, sg);
  }
#pragma unroll 1
  for (int k = 0; k < GS_NS - 1; ++k) {
    if (iu < nunits) {
      cp.issue(sbase + 4u * (uint32_t)((iu % GS_NS) * stage), irow, g.W, g.D, is * S::DC);
      ++iu;
      if (++is == sg.nslab) { is = 0; irow += WD; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- this thread's register tile
  const bool rside = threadIdx.x >= sg.half;
  const int xb = 4 * (rside ? threadIdx.x - sg.half : threadIdx.x);
  const bool active = xb < g.W && (rside ? gr != nullptr : gl != nullptr);
  const float inv = g.mean ? 1.f / (float)CPG : 1.f;   // CPG is a power of two: exact
  Tin* __restrict__ gdst = rside ? gr : gl;
  const bool vec = (g.W & 3) == 0 && (reinterpret_cast<uintptr_t>(gdst) & 15) == 0;
  int unit = 0;

#pragma unroll 1
  for (int64_t r = r0; r < r1; ++r) {
    const int fb = sg.nfb == 2 ? (int)(r - r0) & 1 : 0;
    const float* __restrict__ sL = sF + fb * fsz;
    const float* __restrict__ sR = sL + CPG * sg.FPL;
    float acc[CPG][4];
#pragma unroll
    for (int j = 0; j < CPG; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

#pragma unroll 1
    for (int s = 0; s < sg.nslab; ++s, ++unit) {
      // slab `unit` has landed (for everyone), and everyone is done with slab unit - 1: its stage is refilled
      // (rows of fewer than GS_NS - 1 slabs: the features issued one slab ago must have landed too)
      if (sg.nslab < GS_NS - 1) asm volatile("cp.async.wait_all;" ::: "memory");
      else asm volatile("cp.async.wait_group %0;" ::"n"(GS_NS - 2) : "memory");
      __syncthreads();
      if (s == 0 && sg.nfb == 2 && r + 1 < r1) {       // next row's features into the other buffer
        const int64_t rn = r + 1;
        const int y = (int)(rn % g.H);
        const int64_t t = rn / g.H;
        gs_stage_features<Tin, CPG>(sF + (fb ^ 1) * fsz, L, R, t / g.G, (int)(t % g.G) * CPG, y, g.W, sg);
      }
      if (s == 0 && sg.nfb == 1 && r > r0) {           // one buffer: everyone is done with the previous row's features now
        const int y = (int)(r % g.H);
        const int64_t t = r / g.H;
        gs_stage_features<Tin, CPG>(sF, L, R, t / g.G, (int)(t % g.G) * CPG, y, g.W, sg);
      }
      if (iu < nunits) {
        cp.issue(sbase + 4u * (uint32_t)((iu % GS_NS) * stage), irow, g.W, g.D, is * S::DC);
        ++iu;
        if (++is == sg.nslab) { is = 0; irow += WD; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (s == 0 && sg.nfb == 1 && r > r0) {           // (exposed once per row; only shapes whose features fill the SM)
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
      }

      const int d0 = s * S::DC;
      // left: the slab at this thread's quad, the right features at x - d0 - 4;  right: both at x' + d0
      const float* __restrict__ sGt = sG + (unit % GS_NS) * stage + ((xb + (rside ? d0 : 0)) >> 2) * S::QP;
      const float* __restrict__ frow = rside ? sL + xb + d0 : sR + sg.DPAD + xb - d0 - 4;
      // a left tile only meets the zero margin once d > x + 3; a right tile only zeros once x' + d >= W
      if (active && (rside ? xb + d0 < g.W : d0 <= xb + 3)) {
        const int nq = min(S::DC, g.D - d0) >> 2;      // disparity quads of this slab below D
        constexpr int UQ = S::DC / 4;                  // quads of a whole slab: unrolled (4 or 8 at a time)
        constexpr bool UNROLL = CPG <= (S::F32 ? GS_UNROLL_CPG : GS_UNROLL_CPG / 2);
        if (UNROLL && nq == UQ) {
          if (!rside) gs_left<Tout, CPG, UQ>(acc, sGt, frow, sg.FPR, 0, UQ);
          else gs_right<Tout, CPG, UQ>(acc, sGt, frow, sg.FPL, xb + d0, g.W, 0, UQ);
        } else if (!S::F32 && UNROLL && nq == 4) {   // half a 16-bit slab (D = 48: 32 + 16)
          if (!rside) gs_left<Tout, CPG, 4>(acc, sGt, frow, sg.FPR, 0, 4);
          else gs_right<Tout, CPG, 4>(acc, sGt, frow, sg.FPL, xb + d0, g.W, 0, 4);
        } else {
          if (!rside) gs_left<Tout, CPG, 0>(acc, sGt, frow, sg.FPR, 0, nq);
          else gs_right<Tout, CPG, 0>(acc, sGt, frow, sg.FPL, xb + d0, g.W, 0, nq);
        }
      }
    }

    // ---- scale and store (x contiguous)
    if (active) {
      const int y = (int)(r % g.H);
      const int64_t t = r / g.H;
      const int64_t n = t / g.G;
      const int c0 = (int)(t % g.G) * CPG;
      Tin* __restrict__ dst = gdst + (((int64_t)n * g.C + c0) * g.H + y) * g.W + xb;
#pragma unroll
      for (int j = 0; j < CPG; ++j) {
        Tin* o = dst + (int64_t)j * g.H * g.W;
        if (vec) {
          if constexpr (sizeof(Tin) == 4) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(acc[j][0] * inv, acc[j][1] * inv, acc[j][2] * inv, acc[j][3] * inv));
          } else {
            union { uint2 raw; Tin t4[4]; } pk;
#pragma unroll
            for (int i = 0; i < 4; ++i) pk.t4[i] = from_f<Tin>(acc[j][i] * inv);
            __stcs(reinterpret_cast<uint2*>(o), pk.raw);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (xb + i < g.W) o[i] = from_f<Tin>(acc[j][i] * inv);
        }
      }
    }
  }
}

// Host side: true when the slab kernel was launched (rc holds its status), false when the shape is not covered
// (D % 4 != 0, W > 1024, channels per group not in {1,2,4,8,16}, misaligned gradient, too much shared memory).
template <typename Tin, typename Tout>
static bool launch_groupwise_bwd_slab(const void* gout, const rsm_feat& left, const rsm_feat& right, void* gl, void* gr,
                                      int64_t N, const CorrGeom& g, cudaStream_t st, const char* where, int& rc) {
  using S = GsSlab<Tout>;
  if (g.D <= 0 || g.D % 4 != 0 || !aligned_to(gout, S::PB) || g.W > 2 * GS_MAXT) return false;
  if (g.cpg != 1 && g.cpg != 2 && g.cpg != 4 && g.cpg != 8 && g.cpg != 16) return false;
  GroupSlabGeom sg;
  const int W4 = (g.W + 3) / 4 * 4;
  sg.NQ = W4 / 4 + 2;
  sg.FPL = W4 + 8;
  sg.half = (W4 / 4 + 31) / 32 * 32;
  sg.nslab = (g.D + S::DC - 1) / S::DC;
  sg.DPAD = S::DC * sg.nslab + 4;
  sg.FPR = sg.DPAD + W4;
  sg.rows = N * g.G * (int64_t)g.H;
  constexpr int FEPV = 16 / (int)sizeof(Tin);
  auto vec_ok = [&](const rsm_feat& f) {
    return f.stride_w == 1 && f.stride_n % FEPV == 0 && f.stride_c % FEPV == 0 && f.stride_h % FEPV == 0 && aligned_to(f.data, 16);
  };
  sg.fvec = g.W % FEPV == 0 && vec_ok(left) && vec_ok(right);
  const int slab = sg.NQ * S::QP + (sg.NQ * S::QP & 1), stage = slab + (slab & 2);
  const int ns = GS_STAGES;
  sg.nfb = 2;
  size_t smem = ((size_t)ns * stage + (size_t)2 * g.cpg * (sg.FPL + sg.FPR)) * sizeof(float);
  if (smem > 200 * 1024) {
    sg.nfb = 1;
    smem = ((size_t)ns * stage + (size_t)g.cpg * (sg.FPL + sg.FPR)) * sizeof(float);
  }
  if (smem > 200 * 1024 || sg.rows <= 0 || sg.rows > 2147483647LL) return false;
  const int nt = 2 * sg.half;
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t resident = (int64_t)kNumSMs * per_sm;
    // persistent CTAs: one resident wave, each walking a contiguous run of rows (at least two rows per CTA)
    const int64_t grid = sg.rows < 2 * resident ? (sg.rows + 1) / 2 : resident;
    kern<<<(unsigned)grid, nt, smem, st>>>((const Tout*)gout, view_of(left), view_of(right), (Tin*)gl, (Tin*)gr, g, sg);
    return finish_launch(where);
  };
  // register budget by CTA size: 96 (three CTAs of <= 192 threads per SM; 104 / 112 leave room for two), 128 otherwise
#define RSM_GS_CASE(CPG_)                                                                             \
  case CPG_:                                                                                          \
    rc = nt <= 192 ? launch(groupwise_bwd_slab_kernel<Tin, Tout, CPG_, 192, GS_STAGES>)               \
       : nt <= 256 ? launch(groupwise_bwd_slab_kernel<Tin, Tout, CPG_, 256, GS_STAGES>)               \
                   : launch(groupwise_bwd_slab_kernel<Tin, Tout, CPG_, GS_MAXT, GS_STAGES>);          \
    break;
  switch (g.cpg) {
    RSM_GS_CASE(1) RSM_GS_CASE(2) RSM_GS_CASE(4) RSM_GS_CASE(8)
    default: RSM_GS_CASE(16)
  }
#undef RSM_GS_CASE
  return true;
}

}  // namespace rsm
