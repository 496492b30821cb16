// Library-level entry points of the lanefit_b200 C ABI (include/lanefit_b200.h).
#include <stdlib.h>

#include "lf_common.cuh"

namespace lf {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
static int g_pdl = -1;   // -1: not decided yet (LANEFIT_PDL=1 turns it on; default off, see lf_common.cuh)
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("LANEFIT_PDL");
        g_pdl = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pdl != 0;
}
}  // namespace lf

extern "C" void lf_set_pdl(int on) { lf::g_pdl = on ? 1 : 0; }
extern "C" int lf_get_pdl(void) { return lf::pdl_enabled() ? 1 : 0; }

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
