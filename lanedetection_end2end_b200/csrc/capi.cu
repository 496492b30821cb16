// Library-level entry points of the lanefit_b200 C ABI (include/lanefit_b200.h).
#include "lf_common.cuh"

namespace lf {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
}  // namespace lf

extern "C" int lf_version(void) { return 100; }

extern "C" const char* lf_error_string(int code) {
    switch (code) {
        case LF_OK: return "ok";
        case LF_ERR_INVALID_ARGUMENT: return "invalid argument";
        case LF_ERR_UNSUPPORTED: return "unsupported configuration";
        case LF_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case LF_ERR_CUDA: return "CUDA error (see lf_last_cuda_error)";
        default: return "unknown error";
    }
}

extern "C" const char* lf_last_cuda_error(void) { return cudaGetErrorString(lf::g_last_cuda_error); }
