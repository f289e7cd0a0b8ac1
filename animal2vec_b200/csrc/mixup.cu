// BC-learning waveform mixup (reference nn/data2vec2.py:536-598, compute_gain_torch 453-498).
//   gain_db[b, w] = 10 log10(max(sum_f |rfft(hann * frame_w)|^2 * A(f), 10^(min_db/10)))
//   G[b] = max_w gain_db[b, w];  p = 1 / (1 + 10^((G1 - G2)/20) * (1 - r) / r)
//   out = (p * x[b] + (1 - p) * x[perm[b]]) / sqrt(p^2 + (1 - p)^2)
// The 400-point real DFT is evaluated directly from a twiddle table in shared memory
// (201 bins x 400 taps per frame): tiny next to the rest of the step, no cuFFT.
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

__global__ void __launch_bounds__(256) mixup_gain_kernel(const float* __restrict__ x, const float* __restrict__ hann,
                                                         const float* __restrict__ aweight, int N, int n_fft, int hop,
                                                         int W, float floor_lin, float* __restrict__ gain_db) {
    extern __shared__ float sm[];
    float* frame = sm;              // [n_fft]
    float* tw_c = sm + n_fft;       // [n_fft]
    float* tw_s = tw_c + n_fft;     // [n_fft]
    __shared__ float red[8];
    const int w = blockIdx.x, b = blockIdx.y;
    const float* src = x + (long long)b * N + (long long)w * hop;
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
        frame[i] = src[i] * hann[i];
        float s, c;
        sincospif(2.0f * (float)i / (float)n_fft, &s, &c);
        tw_c[i] = c;
        tw_s[i] = s;
    }
    __syncthreads();
    const int bins = n_fft / 2 + 1;
    float acc = 0.f;
    for (int k = threadIdx.x; k < bins; k += blockDim.x) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < n_fft; ++n) {
            re += frame[n] * tw_c[idx];
            im -= frame[n] * tw_s[idx];
            idx += k;
            if (idx >= n_fft) idx -= n_fft;
        }
        acc += (re * re + im * im) * aweight[k];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float g = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) g += red[i];
        gain_db[(long long)b * W + w] = 10.0f * log10f(fmaxf(g, floor_lin));
    }
}

__global__ void __launch_bounds__(256) mixup_apply_kernel(const float* __restrict__ x, const int* __restrict__ perm,
                                                          const float* __restrict__ gain_db, int N, int W, float r,
                                                          float* __restrict__ out, float* __restrict__ p_out) {
    __shared__ float red[2][8];
    __shared__ float s_p;
    const int b = blockIdx.y;
    const int b2 = perm[b];
    float g1 = -INFINITY, g2 = -INFINITY;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        g1 = fmaxf(g1, gain_db[(long long)b * W + i]);
        g2 = fmaxf(g2, gain_db[(long long)b2 * W + i]);
    }
    g1 = warp_max(g1);
    g2 = warp_max(g2);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = g1;
        red[1][threadIdx.x >> 5] = g2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = -INFINITY, c = -INFINITY;
        for (int i = 0; i < 8; ++i) {
            a = fmaxf(a, red[0][i]);
            c = fmaxf(c, red[1][i]);
        }
        const float p = 1.0f / (1.0f + powf(10.0f, (a - c) / 20.0f) * (1.0f - r) / r);
        s_p = p;
        if (blockIdx.x == 0 && p_out != nullptr) p_out[b] = p;
    }
    __syncthreads();
    const float p = s_p;
    const float inv = rsqrtf(p * p + (1.f - p) * (1.f - p));
    const float* x1 = x + (long long)b * N;
    const float* x2 = x + (long long)b2 * N;
    float* o = out + (long long)b * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        o[i] = (p * x1[i] + (1.f - p) * x2[i]) * inv;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_mixup_gain(const float* x, const float* hann, const float* aweight, int B, int N, int n_fft,
                              int hop, float min_db, float* gain_db, a2v_stream_t stream) {
    A2V_REQUIRE(x && hann && aweight && gain_db, "mixup_gain: NULL pointer");
    A2V_REQUIRE(B > 0 && n_fft >= 2 && n_fft <= 8192 && hop > 0 && N >= n_fft, "mixup_gain: bad extents");
    const int W = (N - n_fft) / hop + 1;
    const size_t smem = (size_t)3 * n_fft * sizeof(float);
    if (smem > 48 * 1024 && a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(mixup_gain_kernel), smem) != A2V_OK)
        return A2V_ERR_CUDA;
    dim3 grid(W, B);
    mixup_gain_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, hann, aweight, N, n_fft, hop, W, powf(10.0f, min_db / 10.0f), gain_db);
    return a2v_check_launch("mixup_gain");
}

extern "C" int a2v_mixup_apply(const float* x, const int32_t* perm, const float* gain_db, int B, int N, int W, float r,
                               float* out, float* p_out, a2v_stream_t stream) {
    A2V_REQUIRE(x && perm && gain_db && out && x != out, "mixup_apply: NULL pointer or in-place call");
    A2V_REQUIRE(B > 0 && N > 0 && W > 0 && r > 0.f && r <= 1.f, "mixup_apply: bad arguments");
    dim3 grid(ceil_div(N, 256 * 8), B);
    mixup_apply_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, perm, gain_db, N, W, r, out, p_out);
    return a2v_check_launch("mixup_apply");
}
