// errors.cu -- thread-local error string and version of libaurdf.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace aurdf {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return AURDF_ECUDA;
}
}  // namespace aurdf

extern "C" int aurdf_version(void) { return AURDF_VERSION; }
extern "C" const char *aurdf_last_error_string(void) { return aurdf::g_err; }
