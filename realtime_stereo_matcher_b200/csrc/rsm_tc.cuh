// tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels of librsm_b200 (sm_100a only):
// the banded inner-product GEMM (rsm_corr_tc.cu) and the v4 per-disparity Conv3d stack as implicit GEMMs
// (rsm_v4vol.cu).  Inline PTX only -- no CUTLASS / CuTe dependency.
#pragma once

#include <cuda.h>   // CUtensorMap (types only; the encoder is looked up at run time, no libcuda link dependency)

#include "rsm_common.cuh"

namespace rsm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor bit layout).
// MN-major operands: SBO = stride between 8-element groups along M/N, LBO = stride between 8-row groups
// along K (verified on B200 against the oracle; the swapped assignment produces garbage).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;   // 0 = SWIZZLE_NONE, 1 = SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- TMA: one box of a (W, H, C, N) tensor map -> shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t mbar, int x, int y, int c, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(mbar), "r"(x), "r"(y), "r"(c), "r"(n)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {   // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// wait for completion of the phase with the given parity; bounded so a protocol bug cannot hang the GPU.
// try_wait carries a suspend-time hint: the waiting thread sleeps in hardware until the phase completes (or the
// hint expires) instead of spinning -- a spinning issuer / producer lane was taking half the issue slots of its
// scheduler away from the epilogue warps that share it (measured: those warps ran 2x slower).
// A wait that expires (4096 x 10 ms) is a protocol failure: EVERY role stops there -- the kernel traps, the launch
// fails with a sticky CUDA error and the next rsm_* call on the device returns RSM_ERR_CUDA.  No role ever runs on
// past a failed wait (it would overwrite a shared-memory stage or a TMEM buffer that is still in use), so the
// kernel cannot return RSM_OK with a poisoned volume.
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
#pragma unroll 1   // (unrolled 16x by default: ~560 instructions per call site)
  for (int it = 0; it < (1 << 12); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no -lcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encoder() {
  static const TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<TmapEncodeFn>(p);
  }();
  return fn;
}

}  // namespace rsm
