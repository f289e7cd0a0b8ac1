// C-ABI plumbing shared by all kernels: thread-local error string, launch check,
// device queries.
#include <stdarg.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <utility>
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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function setting: the opt-in is recorded per
// (function, device) under a mutex, so a process that drives several GPUs (or launches from several threads) configures
// every kernel on every device exactly once and never skips a device because another one was configured first.
int a2v_ensure_dynamic_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        a2v_set_error("cudaGetDevice failed");
        return A2V_ERR_CUDA;
    }
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = done[std::make_pair(func, dev)];
    if (bytes <= have) return A2V_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        a2v_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %zu) failed: %s", bytes, cudaGetErrorString(e));
        return A2V_ERR_CUDA;
    }
    have = bytes;
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
