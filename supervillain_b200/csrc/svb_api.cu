// svb_api.cu -- version, error reporting and small host-side helpers of libsvb200.
#include <stdarg.h>
#include <string.h>

#include "svb_common.cuh"

namespace svb {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_last_error, sizeof(g_last_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

}  // namespace svb

extern "C" int svb_version(void) { return SVB_VERSION_MAJOR * 1000 + SVB_VERSION_MINOR; }

extern "C" const char* svb_last_error(void) { return svb::g_last_error; }

extern "C" void svb_philox4x32_10_host(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    svb::Philox4 p = svb::philox4x32<10>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = p.w;
}
