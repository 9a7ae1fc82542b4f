// Error reporting and version entry points of the C ABI (include/tcsfm.h).
#include <cstdarg>
#include "common.cuh"

namespace tcsfm {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return 2;
    }
    return 0;
}

}  // namespace tcsfm

extern "C" const char* tcsfm_last_error(void) { return tcsfm::g_last_error; }
extern "C" int tcsfm_abi_version(void) { return TCSFM_ABI_VERSION; }
