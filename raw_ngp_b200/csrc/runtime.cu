// runtime.cu -- library identity and error reporting for libngp_b200.
#include "common.cuh"
#include <string.h>

namespace ngp {
static thread_local char g_last_error[256] = "";
void set_last_cuda_error(cudaError_t e) {
    const char* s = cudaGetErrorString(e);
    strncpy(g_last_error, s ? s : "unknown CUDA error", sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = 0;
}
}  // namespace ngp

extern "C" int ngp_abi_version(void) { return NGP_B200_ABI_VERSION; }

extern "C" const char* ngp_status_string(int status) {
    switch (status) {
        case NGP_OK: return "ok";
        case NGP_ERR_BAD_DTYPE: return "unsupported dtype (expected NGP_F32, NGP_F16 or NGP_BF16)";
        case NGP_ERR_UNSUPPORTED: return "unsupported dimension (D in {2,3}, C in {1,2,4,8}, SH degree in 1..8, MLP dims multiple of 16 and <= 128)";
        case NGP_ERR_NULL: return "required pointer is NULL";
        case NGP_ERR_ALIGN: return "pointer is not aligned for the kernel's vector width";
        case NGP_ERR_CUDA: return "CUDA launch failed (see ngp_last_cuda_error)";
        case NGP_ERR_BAD_ARG: return "argument out of range";
        default: return "unknown status";
    }
}

extern "C" const char* ngp_last_cuda_error(void) { return ngp::g_last_error; }
