// MUFU.EX2 / FFMA2 / F2FP issue-rate microbenchmark: lane-ops per clock per SM for 1..8 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[16]; unsigned acc = 0;
    for (int i = 0; i < 16; ++i) x[i] = -1.0f - 0.001f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); if (i & 1) { unsigned pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[i]), "f"(x[i - 1])); x[i] += __uint_as_float(pk & 0x3f800000u); } }
            if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); if (i & 1) { unsigned pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[i]), "f"(x[i - 1])); acc ^= pk; } }
            if (MODE == 5) { if (i & 1) { unsigned pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[i]), "f"(x[i - 1])); acc ^= pk; x[i] = __uint_as_float(acc & 0x3fffffffu); } }
            if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); asm volatile("add.f32 %0, %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    float s = __uint_as_float(acc & 0x3f800000u);
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int mode = 0; mode < 6; ++mode)
        for (int threads = 128; threads <= 1024; threads *= 2) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
            if (mode == 4) k<4><<<148, threads>>>(out, cyc, iters);
            if (mode == 5) k<5><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
            printf("mode %d (%s) warps/scheduler %d: %.2f lane-ops of the first kind per clock per SM\n", mode,
                   mode == 0 ? "ex2 only" : mode == 1 ? "ffma only" : mode == 2 ? "ex2 + 3 fp32 ops" : mode == 3 ? "ex2 + half a bf16x2 pack (dependent)" : mode == 4 ? "ex2 + half a bf16x2 pack (independent)" : "bf16x2 pack only (per 2 floats: x0.5)", threads / 128, (double)threads * 16 * iters / c);
        }
    return 0;
}
