// Channels-last im2col / col2im for the strided feature-extractor convolutions
// (reference nn/utils.py:1085-1090: Conv1d(k, stride s, padding ceil(s/2)), layers 1-4 of
// '[(512,10,5)] + [(512,3,2)]*3'). The column matrix feeds the tcgen05 GEMM; stride-1 layers
// skip this and use the GEMM's tap loop directly.
//   col[b, t', j*C + c] = x[b, t'*s + j - pad, c]   (0 outside [0, Tin))
//   dx[b, t, c]         = sum_{j : (t + pad - j) % s == 0} dcol[b, (t + pad - j)/s, j*C + c]
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, T* __restrict__ col, int B, int Tin,
                                                     int Tout, int C, int k, int s, int pad) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const long long total = (long long)B * Tout * k;
    for (long long o = warp0; o < total; o += nwarps) {
        const int j = (int)(o % k);
        const long long bt = o / k;
        const int tp = (int)(bt % Tout);
        const long long b = bt / Tout;
        const int t = tp * s + j - pad;
        T* dst = col + bt * (long long)(k * C) + (long long)j * C;
        if (t >= 0 && t < Tin) {
            const T* src = x + (b * Tin + t) * C;
            for (int c = lane * 4; c < C; c += 128) {
                float v[4];
                load4(src + c, v);
                store4(dst + c, v);
            }
        } else {
            const float z[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = lane * 4; c < C; c += 128) store4(dst + c, z);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int B, int Tin,
                                                     int Tout, int C, int k, int s, int pad) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const long long total = (long long)B * Tin;
    for (long long o = warp0; o < total; o += nwarps) {
        const int t = (int)(o % Tin);
        const long long b = o / Tin;
        for (int c = lane * 4; c < C; c += 128) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < k; ++j) {
                const int u = t + pad - j;
                if (u < 0 || (u % s) != 0) continue;
                const int tp = u / s;
                if (tp >= Tout) continue;
                float v[4];
                load4(dcol + (b * Tout + tp) * (long long)(k * C) + (long long)j * C + c, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += v[q];
            }
            store4(dx + o * C + c, acc);
        }
    }
}

static int rows_grid(long long rows) {
    long long b = ceil_div64(rows, 8);
    long long cap = (long long)a2v_num_sms() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace a2v

using namespace a2v;

static int check_i2c(int dtype, const void* a, const void* b, int B, int Tin, int Tout, int C, int k, int s, int pad) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "im2col: bad dtype");
    A2V_REQUIRE(a && b, "im2col: NULL pointer");
    A2V_REQUIRE(B > 0 && Tin > 0 && Tout > 0 && C > 0 && C % 4 == 0 && k > 0 && s > 0 && pad >= 0, "im2col: bad extents");
    A2V_REQUIRE((long long)(Tout - 1) * s - pad < Tin, "im2col: last window starts beyond the input");
    return A2V_OK;
}

extern "C" int a2v_im2col(int dtype, const void* x, void* col, int B, int Tin, int Tout, int C, int k, int stride,
                          int pad, a2v_stream_t stream) {
    int rc = check_i2c(dtype, x, col, B, Tin, Tout, C, k, stride, pad);
    if (rc != A2V_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = rows_grid((long long)B * Tout * k);
    if (dtype == A2V_F32)
        im2col_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)col, B, Tin, Tout, C, k, stride, pad);
    else
        im2col_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)col, B, Tin, Tout, C, k, stride, pad);
    return a2v_check_launch("im2col");
}

extern "C" int a2v_col2im(int dtype, const void* dcol, void* dx, int B, int Tin, int Tout, int C, int k, int stride,
                          int pad, a2v_stream_t stream) {
    int rc = check_i2c(dtype, dcol, dx, B, Tin, Tout, C, k, stride, pad);
    if (rc != A2V_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = rows_grid((long long)B * Tin);
    if (dtype == A2V_F32)
        col2im_kernel<float><<<grid, 256, 0, st>>>((const float*)dcol, (float*)dx, B, Tin, Tout, C, k, stride, pad);
    else
        col2im_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dcol, (bf16*)dx, B, Tin, Tout, C, k, stride, pad);
    return a2v_check_launch("col2im");
}
