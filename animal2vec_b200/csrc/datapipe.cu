// Input-pipeline kernels: what the reference does per clip on the host in its Dataset
// (file:line under /root/reference), done for a whole collated batch on the device.
//   fairseq RawAudioDataset.postprocess (called at nn/audio_tasks.py:332; task.normalize = true):
//       feats = F.layer_norm(feats, feats.shape)  per clip, biased variance, eps 1e-5       -> a2v_clip_layer_norm
//   nn/audio_tasks.py:336-381 frame-level multi-hot targets: sample-level label vector (wav_len, classes) from the
//       (start, end, category, focal) intervals, sampled at round(linspace(0, wav_len, T, endpoint=False))
//       (scipy interp1d at integer sample positions = plain indexing)                      -> a2v_frame_labels
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

// one block per clip: two passes over the clip (sum / sum of squares with a first-sample shift, then normalise)
__global__ void __launch_bounds__(1024) clip_layer_norm_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                               float eps) {
    __shared__ float red[2][32];
    __shared__ float stat[2];
    const float* xr = x + (long long)blockIdx.x * N;
    float* yr = y + (long long)blockIdx.x * N;
    const float shift = xr[0];
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x * 4; i < N; i += blockDim.x * 4) {
        if (i + 3 < N && (reinterpret_cast<uintptr_t>(xr + i) & 15) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(xr + i);
            const float d0 = v.x - shift, d1 = v.y - shift, d2 = v.z - shift, d3 = v.w - shift;
            s1 += d0 + d1 + d2 + d3;
            s2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        } else {
            for (int j = i; j < N && j < i + 4; ++j) {
                const float d = xr[j] - shift;
                s1 += d;
                s2 += d * d;
            }
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s1;
        red[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float a = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.f;
        float b = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        if (threadIdx.x == 0) {
            const float m = a / (float)N;
            stat[0] = shift + m;
            stat[1] = rsqrtf(fmaxf(b / (float)N - m * m, 0.f) + eps);
        }
    }
    __syncthreads();
    const float mean = stat[0], rstd = stat[1];
    for (int i = threadIdx.x; i < N; i += blockDim.x) yr[i] = (xr[i] - mean) * rstd;
}

// out[b, t, c] = 1 iff some interval i of clip b (intervals [off[b], off[b+1])) with category c -- or, for the last class
// when focal prediction is on, with focal flag 1 -- covers sample round-half-even(t * wav_len / T).
__global__ void __launch_bounds__(256) frame_labels_kernel(const int* __restrict__ off, const int* __restrict__ start,
                                                           const int* __restrict__ end, const int* __restrict__ cat,
                                                           const int* __restrict__ foc, int B, int T, int C, int wav_len,
                                                           int focal_class, float* __restrict__ out) {
    const long long n = (long long)B * T;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / T), t = (int)(e - (long long)b * T);
        // np.round(np.linspace(0, wav_len, T, endpoint=False)): t * (wav_len / T) in float64, round half to even
        const long long s = (long long)rint((double)t * ((double)wav_len / (double)T));
        float* o = out + e * C;
        for (int c = 0; c < C; ++c) o[c] = 0.f;
        for (int i = off[b]; i < off[b + 1]; ++i) {
            if (s >= start[i] && s < end[i]) {
                if (cat[i] >= 0 && cat[i] < C) o[cat[i]] = 1.f;
                if (focal_class >= 0 && foc != nullptr && foc[i] == 1) o[focal_class] = 1.f;
            }
        }
    }
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_clip_layer_norm(const float* x, float* y, int B, int N, float eps, a2v_stream_t stream) {
    A2V_REQUIRE(x && y && B >= 0 && N > 0, "clip_layer_norm: bad arguments");
    if (B == 0) return A2V_OK;
    clip_layer_norm_kernel<<<B, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, N, eps);
    return a2v_check_launch("clip_layer_norm");
}

extern "C" int a2v_frame_labels(const int32_t* offsets, const int32_t* start, const int32_t* end, const int32_t* cat,
                                const int32_t* foc, int B, int T, int C, int wav_len, int focal_class, float* out,
                                a2v_stream_t stream) {
    A2V_REQUIRE(offsets && out && B >= 0 && T > 0 && C > 0 && wav_len > 0, "frame_labels: bad arguments");
    if (B == 0) return A2V_OK;
    long long blocks = ceil_div64((long long)B * T, 256);
    const long long cap = (long long)a2v_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    frame_labels_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(offsets, start, end, cat, foc, B, T,
                                                                                         C, wav_len, focal_class, out);
    return a2v_check_launch("frame_labels");
}
