// C-ABI plumbing shared by all kernels: thread-local error string, launch check,
// device queries.
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

static thread_local char g_err[512] = "";

void a2v_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int a2v_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        a2v_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return A2V_ERR_CUDA;
    }
    return A2V_OK;
}

extern "C" const char* a2v_last_error(void) { return g_err; }
extern "C" int a2v_version(void) { return 100; }

extern "C" int a2v_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" int a2v_num_sms(void) {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}
