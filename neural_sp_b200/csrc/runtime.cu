// Library-wide plumbing: error strings, device query, version.
#include <stdarg.h>
#include "common.cuh"
#include <stdlib.h>

namespace nsp {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("NSP_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}
}  // namespace nsp

extern "C" int nsp_version(void) { return 100; }
extern "C" const char* nsp_last_error(void) { return nsp::g_err; }

extern "C" nsp_status nsp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0, maj = 0, min = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        nsp::set_error("no CUDA device: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return NSP_ERR_NO_DEVICE;
    }
    NSP_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    NSP_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    NSP_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10) {
        nsp::set_error("device is sm_%d%d; this library is built for sm_100a only", maj, min);
        return NSP_ERR_NO_DEVICE;
    }
    return NSP_OK;
}
