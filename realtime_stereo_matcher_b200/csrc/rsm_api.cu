// Library-level entry points of librsm_b200: version and error reporting.
#include <stdio.h>
#include <string.h>

#include "rsm_common.cuh"

namespace rsm {

static thread_local char g_cuda_detail[256] = {0};

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_detail, sizeof(g_cuda_detail), "CUDA failure in %s: %s (%s)", where,
           cudaGetErrorName(e), cudaGetErrorString(e));
}

static thread_local char g_io_detail[512] = {0};

void set_io_error(const char* what, const char* path) {
  snprintf(g_io_detail, sizeof(g_io_detail), "%s: %s", what, path ? path : "(null)");
}

}  // namespace rsm

extern "C" int rsm_version(void) { return RSM_VERSION; }

extern "C" const char* rsm_last_error(int code) {
  switch (code) {
    case RSM_OK: return "ok";
    case RSM_ERR_INVALID_SHAPE: return "invalid shape (negative size, C % G != 0, or index overflow)";
    case RSM_ERR_UNSUPPORTED_DTYPE: return "unsupported dtype (expected RSM_F32, RSM_F16 or RSM_BF16)";
    case RSM_ERR_NULL_POINTER: return "null pointer for a required tensor";
    case RSM_ERR_CUDA: return rsm::g_cuda_detail[0] ? rsm::g_cuda_detail : "CUDA failure";
    case RSM_ERR_MISALIGNED: return "dense tensor pointer is not aligned to its element size";
    case RSM_ERR_UNSUPPORTED_CONFIG: return "unsupported configuration for this kernel";
    case RSM_ERR_IO: return rsm::g_io_detail[0] ? rsm::g_io_detail : "file input / output failure";
    default: return "unknown rsm status code";
  }
}
