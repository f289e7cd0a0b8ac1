// Teacher-target construction and masked regression loss (HBM-bound).
//
// Reference behaviour replaced (file:line under /root/reference):
//   nn/data2vec2.py:1023-1066 make_targets: per layer F.instance_norm over time (fp32, biased
//     variance, eps 1e-5, no affine) of the top-K teacher FFN outputs, then the mean over layers
//       -> a2v_target_stats + a2v_target_apply
//   data2vec2.py:850-862,1005-1021: y.repeat_interleave(M)[mask], x[mask], MSE * D^-0.5
//       -> a2v_d2v_loss_fwd / _bwd  (the cloned / gathered tensors are never materialised)
//   data2vec2.py:1095-1110 compute_var (pred_var / target_var logging statistics)
//       -> column sums accumulated in the same pass as the loss
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

// stats[l][b][c] = (mean, rstd) over t of layer l. grid (D/128, B, K), 256 threads.
template <typename T>
__global__ void __launch_bounds__(256) target_stats_kernel(const void* const* __restrict__ layers,
                                                           float2* __restrict__ stats, int B, int T_, int D,
                                                           float eps) {
    __shared__ float part[2][8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + lane * 4;
    const int b = blockIdx.y, l = blockIdx.z;
    const T* x = reinterpret_cast<const T*>(layers[l]) + (long long)b * T_ * D;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, shift[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < D) {
        load4(x + c, shift);  // shifted sums: robust against mean^2 >> var cancellation
        for (int t = warp; t < T_; t += 8) {
            float v[4];
            load4(x + (long long)t * D + c, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = v[j] - shift[j];
                s1[j] += d;
                s2[j] += d * d;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        part[0][warp][lane * 4 + j] = s1[j];
        part[1][warp][lane * 4 + j] = s2[j];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cc = blockIdx.x * 128 + threadIdx.x;
        if (cc < D) {
            float a = 0.f, q = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                a += part[0][w][threadIdx.x];
                q += part[1][w][threadIdx.x];
            }
            const float sh = to_f32(x[cc]);
            const float m1 = a / (float)T_;
            const float var = fmaxf(q / (float)T_ - m1 * m1, 0.f);
            stats[((long long)l * B + b) * D + cc] = make_float2(sh + m1, rsqrtf(var + eps));
        }
    }
}

// y[b][t][c] = (1/K) sum_l (x_l[b][t][c] - mean) * rstd.  grid (D/128, B, row_splits)
template <typename T>
__global__ void __launch_bounds__(256) target_apply_kernel(const void* const* __restrict__ layers,
                                                           const float2* __restrict__ stats, float* __restrict__ y,
                                                           int K, int B, int T_, int D, int rows_per_block) {
    extern __shared__ float2 sst[];  // [K][128]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < K * 128; i += blockDim.x) {
        const int l = i / 128, cc = c0 + (i % 128);
        sst[i] = cc < D ? stats[((long long)l * B + b) * D + cc] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int c = c0 + lane * 4;
    if (c >= D) return;
    const int t0 = blockIdx.z * rows_per_block;
    const int t1 = min(T_, t0 + rows_per_block);
    const float invk = 1.0f / (float)K;
    for (int t = t0 + warp; t < t1; t += 8) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const long long off = ((long long)b * T_ + t) * D + c;
        for (int l = 0; l < K; ++l) {
            float v[4];
            load4(reinterpret_cast<const T*>(layers[l]) + off, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 st = sst[l * 128 + lane * 4 + j];
                acc[j] += (v[j] - st.x) * st.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] *= invk;
        store4(y + off, acc);
    }
}

// masked regression loss + logging statistics. grid (D/128, nblk)
template <typename T>
__global__ void __launch_bounds__(256) d2v_loss_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ y,
                                                           const uint8_t* __restrict__ mask, long long RT, int T_,
                                                           int M, int D, float scale, long long rows_per_block,
                                                           double* __restrict__ loss_sum,
                                                           double* __restrict__ colstats) {
    __shared__ float part[5][8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + lane * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > RT) r1 = RT;
    float sx[4] = {0, 0, 0, 0}, sxx[4] = {0, 0, 0, 0}, sy[4] = {0, 0, 0, 0}, syy[4] = {0, 0, 0, 0};
    float loss = 0.f;
    if (c < D) {
        for (long long o = r0 + warp; o < r1; o += 8) {
            if (mask[o] == 0) continue;
            const long long r = o / T_;
            const int t = (int)(o - r * T_);
            float x[4], yy[4];
            load4(pred + o * D + c, x);
            load4(y + ((r / M) * T_ + t) * D + c, yy);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = x[j] - yy[j];
                loss += d * d;
                sx[j] += x[j];
                sxx[j] += x[j] * x[j];
                sy[j] += yy[j];
                syy[j] += yy[j] * yy[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        part[0][warp][lane * 4 + j] = sx[j];
        part[1][warp][lane * 4 + j] = sxx[j];
        part[2][warp][lane * 4 + j] = sy[j];
        part[3][warp][lane * 4 + j] = syy[j];
    }
    loss = warp_sum(loss);
    if (lane == 0) part[4][warp][0] = loss;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cc = blockIdx.x * 128 + threadIdx.x;
        if (cc < D) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += part[k][w][threadIdx.x];
                atomicAdd(colstats + (long long)k * D + cc, (double)s);
            }
        }
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[4][w][0];
        atomicAdd(loss_sum, (double)s * (double)scale);
    }
}

// dpred = mask ? 2 * scale * g * (pred - y) : 0
template <typename T>
__global__ void __launch_bounds__(256) d2v_loss_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ y,
                                                           const uint8_t* __restrict__ mask, T* __restrict__ dpred,
                                                           long long RT, int T_, int M, int D, float coef,
                                                           const float* __restrict__ gptr) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const float g = coef * (gptr != nullptr ? *gptr : 1.0f);
    for (long long o = warp0; o < RT; o += nwarps) {
        const bool m = mask[o] != 0;
        const long long r = o / T_;
        const int t = (int)(o - r * T_);
        for (int c = lane * 4; c < D; c += 128) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m) {
                float x[4], yy[4];
                load4(pred + o * D + c, x);
                load4(y + ((r / M) * T_ + t) * D + c, yy);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = g * (x[j] - yy[j]);
            }
            store4(dpred + o * D + c, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// bf16 fast paths (D % 256 == 0): 16-byte accesses, several independent rows in flight per warp
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// Masked regression loss, its logging statistics AND (optionally) its gradient in one pass over the predictions.
// A warp owns one (clip b, frame t) at a time for a 256-column slab: the fp32 target row segment is read ONCE and
// serves all M clones (the reference materialises y.repeat_interleave(M), nn/data2vec2.py:850-858); the M mask
// bytes are fetched first, then every masked clone's 16-byte prediction chunk is in flight together. With
// WRITE_GRAD the gradient 2 * scale * (x - y) (zeros for unmasked rows) is stored to dpred, which may alias pred.
// grid (D / 256, blocks over (b, t)).
template <bool WRITE_GRAD, int MC>
__global__ void __launch_bounds__(256) d2v_loss_fused_bf16_kernel(const bf16* pred,
                                                                  const float* __restrict__ y,
                                                                  const uint8_t* __restrict__ mask, bf16* dpred,
                                                                  long long BT, int T_, int M, int D, float scale,
                                                                  float gcoef, double* __restrict__ loss_sum,
                                                                  double* __restrict__ colstats) {
    __shared__ float part[4][8][256];
    __shared__ float lpart[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 256 + lane * 8;
    float sx[8], sxx[8], sy[8], syy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sx[j] = sxx[j] = sy[j] = syy[j] = 0.f;
    float loss = 0.f;
    const long long stride = (long long)gridDim.y * 8;
    for (long long bt = (long long)blockIdx.y * 8 + warp; bt < BT; bt += stride) {
        const long long b = bt / T_;
        const int t = (int)(bt - b * T_);
        const float4 y0 = *reinterpret_cast<const float4*>(y + bt * D + c);
        const float4 y1 = *reinterpret_cast<const float4*>(y + bt * D + c + 4);
        const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
        for (int m0 = 0; m0 < M; m0 += MC) {
            bool mk[MC];
            uint4 xv[MC];
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                const int m = m0 + i;
                mk[i] = m < M && mask[(b * M + m) * T_ + t] != 0;
            }
#pragma unroll
            for (int i = 0; i < MC; ++i)
                if (mk[i]) xv[i] = *reinterpret_cast<const uint4*>(pred + ((b * M + m0 + i) * T_ + t) * D + c);
            int nm = 0;
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (m0 + i >= M) continue;
                uint4 gv = make_uint4(0u, 0u, 0u, 0u);
                if (mk[i]) {
                    float x[8], g[8];
                    unpack8(xv[i], x);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d = x[j] - yy[j];
                        loss = fmaf(d, d, loss);
                        sx[j] += x[j];
                        sxx[j] = fmaf(x[j], x[j], sxx[j]);
                        g[j] = gcoef * d;
                    }
                    ++nm;
                    gv.x = pack_bf16x2(g[0], g[1]); gv.y = pack_bf16x2(g[2], g[3]);
                    gv.z = pack_bf16x2(g[4], g[5]); gv.w = pack_bf16x2(g[6], g[7]);
                }
                if (WRITE_GRAD) *reinterpret_cast<uint4*>(dpred + ((b * M + m0 + i) * T_ + t) * D + c) = gv;
            }
            const float fn = (float)nm;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sy[j] = fmaf(fn, yy[j], sy[j]);
                syy[j] = fmaf(fn * yy[j], yy[j], syy[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        part[0][warp][lane * 8 + j] = sx[j];
        part[1][warp][lane * 8 + j] = sxx[j];
        part[2][warp][lane * 8 + j] = sy[j];
        part[3][warp][lane * 8 + j] = syy[j];
    }
    loss = warp_sum(loss);
    if (lane == 0) lpart[warp] = loss;
    __syncthreads();
    {
        const int cc = blockIdx.x * 256 + threadIdx.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += part[k][w][threadIdx.x];
            atomicAdd(colstats + (long long)k * D + cc, (double)s);
        }
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += lpart[w];
        atomicAdd(loss_sum, (double)s * (double)scale);
    }
}

// bf16 instance-norm statistics: 256-column slab per block, four frames in flight per warp. grid (D/256, B, K)
__global__ void __launch_bounds__(256) target_stats_bf16_wide_kernel(const void* const* __restrict__ layers,
                                                                     float2* __restrict__ stats, int B, int T_, int D,
                                                                     float eps) {
    __shared__ float part[2][8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 256 + lane * 8;
    const int b = blockIdx.y, l = blockIdx.z;
    const bf16* x = reinterpret_cast<const bf16*>(layers[l]) + (long long)b * T_ * D;
    float s1[8], s2[8], shift[8];
    unpack8(*reinterpret_cast<const uint4*>(x + c), shift);  // shifted sums: robust against mean^2 >> var
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    for (int t = warp; t < T_; t += 32) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (t + 8 * u < T_) v[u] = *reinterpret_cast<const uint4*>(x + (long long)(t + 8 * u) * D + c);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (t + 8 * u >= T_) continue;
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = f[j] - shift[j];
                s1[j] += d;
                s2[j] = fmaf(d, d, s2[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        part[0][warp][lane * 8 + j] = s1[j];
        part[1][warp][lane * 8 + j] = s2[j];
    }
    __syncthreads();
    const int cc = blockIdx.x * 256 + threadIdx.x;
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        a += part[0][w][threadIdx.x];
        q += part[1][w][threadIdx.x];
    }
    const float sh = to_f32(x[cc]);
    const float m1 = a / (float)T_;
    const float var = fmaxf(q / (float)T_ - m1 * m1, 0.f);
    stats[((long long)l * B + b) * D + cc] = make_float2(sh + m1, rsqrtf(var + eps));
}

// y = (1/K) sum_l (x_l - mean_l) * rstd_l = sum_l x_l * a_l - c  with a_l = rstd_l / K, c = sum_l mean_l * a_l held in
// registers per lane (8 columns); the K layer chunks of a frame are loaded together. grid (D/256, B, row splits)
template <int KC>
__global__ void __launch_bounds__(256) target_apply_bf16_wide_kernel(const void* const* __restrict__ layers,
                                                                     const float2* __restrict__ stats,
                                                                     float* __restrict__ y, int K, int B, int T_, int D,
                                                                     int rows_per_block) {
    extern __shared__ float sa[];  // [K][256] scale a_l
    __shared__ float sc[256];      // offset c
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256;
    const int b = blockIdx.y;
    {
        float off = 0.f;
        const float invk = 1.0f / (float)K;
        for (int l = 0; l < K; ++l) {
            const float2 st = stats[((long long)l * B + b) * D + c0 + threadIdx.x];
            const float a = st.y * invk;
            sa[l * 256 + threadIdx.x] = a;
            off = fmaf(st.x, a, off);
        }
        sc[threadIdx.x] = off;
    }
    __syncthreads();
    const int c = c0 + lane * 8;
    const int t0 = blockIdx.z * rows_per_block;
    const int t1 = min(T_, t0 + rows_per_block);
    for (int t = t0 + warp; t < t1; t += 8) {
        const long long off = ((long long)b * T_ + t) * D + c;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = -sc[lane * 8 + j];
        for (int l0 = 0; l0 < K; l0 += KC) {
            uint4 v[KC];
#pragma unroll
            for (int i = 0; i < KC; ++i)
                if (l0 + i < K) v[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(layers[l0 + i]) + off);
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                if (l0 + i >= K) continue;
                float f[8];
                unpack8(v[i], f);
                const float4 a0 = *reinterpret_cast<const float4*>(sa + (l0 + i) * 256 + lane * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(sa + (l0 + i) * 256 + lane * 8 + 4);
                acc[0] = fmaf(f[0], a0.x, acc[0]); acc[1] = fmaf(f[1], a0.y, acc[1]);
                acc[2] = fmaf(f[2], a0.z, acc[2]); acc[3] = fmaf(f[3], a0.w, acc[3]);
                acc[4] = fmaf(f[4], a1.x, acc[4]); acc[5] = fmaf(f[5], a1.y, acc[5]);
                acc[6] = fmaf(f[6], a1.z, acc[6]); acc[7] = fmaf(f[7], a1.w, acc[7]);
            }
        }
        *reinterpret_cast<float4*>(y + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(y + off + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_target_stats(int dtype, const void* const* layers_dev, int K, int B, int T, int D, float eps,
                                float* stats, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "target_stats: bad dtype");
    A2V_REQUIRE(layers_dev && stats && K > 0 && B > 0 && T > 0 && D > 0 && D % 4 == 0, "target_stats: bad arguments");
    dim3 grid(ceil_div(D, 128), B, K);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_F32)
        target_stats_kernel<float><<<grid, 256, 0, st>>>(layers_dev, (float2*)stats, B, T, D, eps);
    else if (D % 256 == 0)
        target_stats_bf16_wide_kernel<<<dim3(D / 256, B, K), 256, 0, st>>>(layers_dev, (float2*)stats, B, T, D, eps);
    else
        target_stats_kernel<bf16><<<grid, 256, 0, st>>>(layers_dev, (float2*)stats, B, T, D, eps);
    return a2v_check_launch("target_stats");
}

extern "C" int a2v_target_apply(int dtype, const void* const* layers_dev, int K, int B, int T, int D,
                                const float* stats, float* y, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "target_apply: bad dtype");
    A2V_REQUIRE(layers_dev && stats && y && K > 0 && K <= 64 && B > 0 && T > 0 && D > 0 && D % 4 == 0,
                "target_apply: bad arguments (K <= 64)");
    const int cblocks = ceil_div(D, 128);
    int splits = (a2v_num_sms() * 4) / (cblocks * B);
    if (splits < 1) splits = 1;
    if (splits > ceil_div(T, 8)) splits = ceil_div(T, 8);
    const int rpb = ceil_div(T, splits);
    dim3 grid(cblocks, B, ceil_div(T, rpb));
    const size_t smem = (size_t)K * 128 * sizeof(float2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_BF16 && D % 256 == 0) {
        const int cb = D / 256;
        int sp = (a2v_num_sms() * 4) / (cb * B);
        if (sp < 1) sp = 1;
        if (sp > ceil_div(T, 8)) sp = ceil_div(T, 8);
        const int rp = ceil_div(T, sp);
        target_apply_bf16_wide_kernel<8><<<dim3(cb, B, ceil_div(T, rp)), 256, (size_t)K * 256 * sizeof(float), st>>>(
            layers_dev, (const float2*)stats, y, K, B, T, D, rp);
        return a2v_check_launch("target_apply");
    }
    if (dtype == A2V_F32)
        target_apply_kernel<float><<<grid, 256, smem, st>>>(layers_dev, (const float2*)stats, y, K, B, T, D, rpb);
    else
        target_apply_kernel<bf16><<<grid, 256, smem, st>>>(layers_dev, (const float2*)stats, y, K, B, T, D, rpb);
    return a2v_check_launch("target_apply");
}

extern "C" int a2v_d2v_loss_fused(int dtype, const void* pred, const float* y, const uint8_t* mask, void* dpred,
                                  int64_t R, int T, int clones, int D, float scale, float grad_coef, double* loss_sum,
                                  double* colstats, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_BF16, "d2v_loss_fused: bf16 only (fp32 validation mode uses the separate kernels)");
    A2V_REQUIRE(pred && y && mask && loss_sum && colstats && R >= 0 && T > 0 && clones >= 1 && clones <= 64,
                "d2v_loss_fused: bad arguments");
    A2V_REQUIRE(D > 0 && D % 256 == 0 && R % clones == 0, "d2v_loss_fused: D %% 256 == 0 and R %% clones == 0 required");
    A2V_REQUIRE(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dpred)) & 15) == 0,
                "d2v_loss_fused: pred / y / dpred must be 16-byte aligned");
    if (R == 0) return A2V_OK;
    const long long BT = (long long)(R / clones) * T;
    const int cb = D / 256;
    long long nblk = (long long)a2v_num_sms() * 6 / cb;
    if (nblk < 1) nblk = 1;
    if (nblk > ceil_div64(BT, 8)) nblk = ceil_div64(BT, 8);
    dim3 grid(cb, (unsigned)nblk);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dpred != nullptr)
        d2v_loss_fused_bf16_kernel<true, 6><<<grid, 256, 0, st>>>((const bf16*)pred, y, mask, (bf16*)dpred, BT, T, clones, D,
                                                                  scale, grad_coef, loss_sum, colstats);
    else
        d2v_loss_fused_bf16_kernel<false, 6><<<grid, 256, 0, st>>>((const bf16*)pred, y, mask, nullptr, BT, T, clones, D,
                                                                   scale, 0.f, loss_sum, colstats);
    return a2v_check_launch("d2v_loss_fused");
}

extern "C" int a2v_d2v_loss_fwd(int dtype, const void* pred, const float* y, const uint8_t* mask, int64_t R, int T,
                                int clones, int D, float scale, double* loss_sum, double* colstats,
                                a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "d2v_loss_fwd: bad dtype");
    A2V_REQUIRE(pred && y && mask && loss_sum && colstats && R >= 0 && T > 0 && clones >= 1 && D > 0 && D % 4 == 0,
                "d2v_loss_fwd: bad arguments");
    if (R == 0) return A2V_OK;
    if (dtype == A2V_BF16 && D % 256 == 0 && R % clones == 0)
        return a2v_d2v_loss_fused(dtype, pred, y, mask, nullptr, R, T, clones, D, scale, 0.f, loss_sum, colstats, stream);
    const long long RT = (long long)R * T;
    const int cblocks = ceil_div(D, 128);
    long long nblk = (long long)a2v_num_sms() * 8 / cblocks;
    if (nblk < 1) nblk = 1;
    if (nblk > ceil_div64(RT, 8)) nblk = ceil_div64(RT, 8);
    const long long rpb = ceil_div64(RT, nblk);
    dim3 grid(cblocks, (unsigned)ceil_div64(RT, rpb));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_F32)
        d2v_loss_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, y, mask, RT, T, clones, D, scale, rpb,
                                                          loss_sum, colstats);
    else
        d2v_loss_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)pred, y, mask, RT, T, clones, D, scale, rpb,
                                                         loss_sum, colstats);
    return a2v_check_launch("d2v_loss_fwd");
}

extern "C" int a2v_d2v_loss_bwd(int dtype, const void* pred, const float* y, const uint8_t* mask, void* dpred,
                                int64_t R, int T, int clones, int D, float scale, const float* grad_out_dev,
                                a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "d2v_loss_bwd: bad dtype");
    A2V_REQUIRE(pred && y && mask && dpred && R >= 0 && T > 0 && clones >= 1 && D > 0 && D % 4 == 0,
                "d2v_loss_bwd: bad arguments");
    if (R == 0) return A2V_OK;
    const long long RT = (long long)R * T;
    long long blocks = ceil_div64(RT, 8);
    const long long cap = (long long)a2v_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_F32)
        d2v_loss_bwd_kernel<float><<<(int)blocks, 256, 0, st>>>((const float*)pred, y, mask, (float*)dpred, RT, T,
                                                                 clones, D, 2.f * scale, grad_out_dev);
    else
        d2v_loss_bwd_kernel<bf16><<<(int)blocks, 256, 0, st>>>((const bf16*)pred, y, mask, (bf16*)dpred, RT, T, clones,
                                                                D, 2.f * scale, grad_out_dev);
    return a2v_check_launch("d2v_loss_bwd");
}
