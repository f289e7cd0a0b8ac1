// Finetune head and criterion kernels (BASELINE configs[3]).
//
// Reference behaviour replaced (file:line under /root/reference):
//   nn/wav2vec2.py:446-464   x = mean of the top-k FFN outputs of the main blocks; final dropout (0 in the shipped
//                            recipe); proj = Linear(D, classes)              -> a2v_layer_mean_head_fwd / a2v_head_bwd
//   nn/wav2vec2.py:424-431   target mixup  t' = r * t + (1 - r) * t[perm]    -> folded into the loss kernels
//   nn/utils.py:971-1010     sigmoid_focal_loss (alpha 0.25, gamma 2, fp32)  -> a2v_focal_loss_fwd / _bwd
//   nn/criterions.py:198-229 compute_accuracy / compute_prec_rec_f1 + nn/utils.py:925-969 confusion (multi-label
//                            branch: sigmoid >= threshold against the int64-truncated target, micro-summed)
//                                                                           -> counters of a2v_focal_loss_fwd
//   nn/modalities/base.py:470-484 channel masking (index_put(x, mask_channel, 0)) -> a2v_channel_mask
// All memory-bound row kernels: one warp per row for the head (K * D * 2 bytes read per row), 16-byte accesses.
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int HEAD_CMAX = 32;  // classes

// logits[r, c] = bias[c] + sum_d mean_k(layer_k[r, d]) * W[c, d];  xmean[r, :] optionally stored for the backward
template <typename T>
__global__ void __launch_bounds__(256) layer_mean_head_fwd_kernel(const T* const* __restrict__ layers, int K,
                                                                   long long rows, int D, int C,
                                                                   const float* __restrict__ W,
                                                                   const float* __restrict__ bias, T* __restrict__ xmean,
                                                                   float* __restrict__ logits) {
    extern __shared__ float sW[];  // (C, D)
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const float inv_k = 1.0f / (float)K;
    for (long long r = warp0; r < rows; r += nwarps) {
        float acc[HEAD_CMAX];
#pragma unroll
        for (int c = 0; c < HEAD_CMAX; ++c) acc[c] = 0.f;
        for (int col = lane * 4; col < D; col += 128) {
            float m[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < K; ++k) {
                float v[4];
                load4(layers[k] + r * D + col, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] *= inv_k;
            if (xmean != nullptr) store4(xmean + r * D + col, m);
            if (sizeof(T) == 2) {  // the head multiplies what the backward will see: the stored (rounded) mean
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = to_f32(from_f32<T>(m[j]));
            }
#pragma unroll
            for (int c = 0; c < HEAD_CMAX; ++c) {
                if (c < C) {
                    const float4 w = *reinterpret_cast<const float4*>(sW + c * D + col);
                    acc[c] = fmaf(m[0], w.x, fmaf(m[1], w.y, fmaf(m[2], w.z, fmaf(m[3], w.w, acc[c]))));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HEAD_CMAX; ++c) {
            if (c < C) {
                const float s = warp_sum(acc[c]);
                if (lane == 0) logits[r * C + c] = s + (bias != nullptr ? bias[c] : 0.f);
            }
        }
    }
}

// g[r, :] = (1/K) * sum_c dlogits[r, c] * W[c, :]   (gradient every averaged layer output receives; optional)
// dW[c, :] += sum_r dlogits[r, c] * xmean[r, :],  db[c] += sum_r dlogits[r, c]
// Thread t owns columns 4t .. 4t+3 of a 1024-column slab (grid.y slabs), a block walks a strided set of rows.
template <typename T, int CT>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ xmean,
                                                        const float* __restrict__ W, int K, long long rows, int D, int C,
                                                        T* __restrict__ g, float* __restrict__ dW,
                                                        float* __restrict__ db) {
    const int col = blockIdx.y * 1024 + threadIdx.x * 4;
    const bool col_ok = col < D;
    float w[CT][4], acc[CT][4];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[c][j] = 0.f;
            w[c][j] = (c < C && col_ok) ? W[c * D + col + j] : 0.f;
        }
    }
    float dbacc = 0.f;
    const float inv_k = 1.0f / (float)K;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        float dl[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) dl[c] = c < C ? dlogits[r * C + c] : 0.f;
        if (blockIdx.y == 0 && (int)threadIdx.x < C) dbacc += dlogits[r * C + threadIdx.x];
        if (!col_ok) continue;
        float x[4];
        load4(xmean + r * D + col, x);
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CT; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[c][j] = fmaf(dl[c], x[j], acc[c][j]);
                gv[j] = fmaf(dl[c], w[c][j], gv[j]);
            }
        }
        if (g != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) gv[j] *= inv_k;
            store4(g + r * D + col, gv);
        }
    }
    if (col_ok) {
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (c < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(dW + c * D + col + j, acc[c][j]);
            }
    }
    if (db != nullptr && blockIdx.y == 0 && (int)threadIdx.x < C) atomicAdd(db + threadIdx.x, dbacc);
}

struct FocalParams {
    const float* logits;
    const float* targets;
    const int* perm;  // optional target mixup partner per clip
    long long rows;
    int C;
    int rows_per_clip;
    float r, alpha, gamma, threshold;
};

__device__ __forceinline__ float focal_target(const FocalParams& p, long long row, int c) {
    float t = p.targets[row * p.C + c];
    if (p.perm != nullptr) {
        const long long clip = row / p.rows_per_clip;
        const long long prow = (long long)p.perm[clip] * p.rows_per_clip + (row - clip * p.rows_per_clip);
        t = t * p.r + (1.0f - p.r) * p.targets[prow * p.C + c];
    }
    return t;
}

// counters: [tp, fp, tn, fn, n_correct] (nn/utils.py:925-969 on the int64-truncated target; criterions.py:198-216)
__global__ void __launch_bounds__(256) focal_loss_fwd_kernel(const FocalParams p, double* __restrict__ loss_sum,
                                                             float* __restrict__ loss_out, float* __restrict__ mixed_targets,
                                                             unsigned long long* __restrict__ counters) {
    __shared__ float s_loss[8];
    __shared__ unsigned int s_cnt[5];
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const long long n = p.rows * p.C;
    float lsum = 0.f;
    unsigned int cnt[5] = {0u, 0u, 0u, 0u, 0u};
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / p.C;
        const int c = (int)(e - row * p.C);
        const float x = p.logits[e];
        const float t = focal_target(p, row, c);
        const float pr = 1.0f / (1.0f + expf(-x));
        // binary_cross_entropy_with_logits: max(x, 0) - x t + log(1 + exp(-|x|))
        const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        const float p_t = pr * t + (1.0f - pr) * (1.0f - t);
        float loss = ce * powf(1.0f - p_t, p.gamma);
        if (p.alpha >= 0.f) loss *= p.alpha * t + (1.0f - p.alpha) * (1.0f - t);
        lsum += loss;
        if (loss_out != nullptr) loss_out[e] = loss;
        if (mixed_targets != nullptr) mixed_targets[e] = t;
        if (counters != nullptr) {
            const bool pred = !(pr < p.threshold);  // torch.where(preds < thr, 0, 1)
            const long long ti = (long long)t;      // target.to(torch.int64)
            if (pred && ti == 1) ++cnt[0];
            else if (pred && ti == 0) ++cnt[1];
            else if (!pred && ti == 0) ++cnt[2];
            else if (!pred && ti != 0) ++cnt[3];
            if ((long long)(pred ? 1 : 0) == ti) ++cnt[4];
        }
    }
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = lsum;
    if (counters != nullptr) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            unsigned int v = cnt[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_loss[w];
        atomicAdd(loss_sum, (double)s);
    }
    if (counters != nullptr && threadIdx.x < 5 && s_cnt[threadIdx.x])
        atomicAdd(counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// d loss / d logit = alpha_t * [ (p - t) (1 - p_t)^gamma - ce * gamma (1 - p_t)^(gamma - 1) * p (1 - p) (2 t - 1) ]
__global__ void __launch_bounds__(256) focal_loss_bwd_kernel(const FocalParams p, const float* __restrict__ grad_out,
                                                             float* __restrict__ dlogits) {
    const float go = grad_out != nullptr ? *grad_out : 1.0f;
    const long long n = p.rows * p.C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / p.C;
        const int c = (int)(e - row * p.C);
        const float x = p.logits[e];
        const float t = focal_target(p, row, c);
        const float pr = 1.0f / (1.0f + expf(-x));
        const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        const float p_t = pr * t + (1.0f - pr) * (1.0f - t);
        const float om = 1.0f - p_t;
        const float mod = powf(om, p.gamma);
        const float dmod = om > 0.f ? p.gamma * powf(om, p.gamma - 1.0f) : 0.f;
        float d = (pr - t) * mod - ce * dmod * pr * (1.0f - pr) * (2.0f * t - 1.0f);
        if (p.alpha >= 0.f) d *= p.alpha * t + (1.0f - p.alpha) * (1.0f - t);
        dlogits[e] = d * go;
    }
}

// x[b, t, c] = 0 where chmask[b, c] != 0 (in place)
template <typename T>
__global__ void __launch_bounds__(256) channel_mask_kernel(T* __restrict__ x, const uint8_t* __restrict__ chmask,
                                                           long long rows, int rows_per_clip, int D) {
    const long long n4 = rows * (D / 4);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / (D / 4);
        const int c = (int)(e - row * (D / 4)) * 4;
        const uint32_t m = *reinterpret_cast<const uint32_t*>(chmask + (row / rows_per_clip) * D + c);
        if (m == 0u) continue;
        float v[4];
        load4(x + row * D + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((m >> (8 * j)) & 0xffu) v[j] = 0.f;
        store4(x + row * D + c, v);
    }
}

static int flat_blocks(long long n, int per_block) {
    long long b = ceil_div64(n, per_block);
    const long long cap = (long long)a2v_num_sms() * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_layer_mean_head_fwd(int dtype, const void* const* layers, int K, int64_t rows, int D, int C,
                                       const float* W, const float* bias, void* xmean, float* logits,
                                       a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "layer_mean_head_fwd: bad dtype");
    A2V_REQUIRE(layers && W && logits && K >= 1 && rows >= 0, "layer_mean_head_fwd: bad arguments");
    A2V_REQUIRE(D > 0 && D % 128 == 0 && C >= 1 && C <= HEAD_CMAX, "layer_mean_head_fwd: D %% 128 == 0 and 1 <= classes <= %d",
                HEAD_CMAX);
    if (rows == 0) return A2V_OK;
    const size_t smem = (size_t)C * D * sizeof(float);
    A2V_REQUIRE(smem <= 200 * 1024, "layer_mean_head_fwd: head weight (%zu bytes) does not fit shared memory", smem);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_blocks(rows, 8);
    if (dtype == A2V_F32) {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(layer_mean_head_fwd_kernel<float>), smem) != A2V_OK)
            return A2V_ERR_CUDA;
        layer_mean_head_fwd_kernel<float><<<grid, 256, smem, st>>>(reinterpret_cast<const float* const*>(layers), K, rows, D, C,
                                                                    W, bias, reinterpret_cast<float*>(xmean), logits);
    } else {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(layer_mean_head_fwd_kernel<bf16>), smem) != A2V_OK)
            return A2V_ERR_CUDA;
        layer_mean_head_fwd_kernel<bf16><<<grid, 256, smem, st>>>(reinterpret_cast<const bf16* const*>(layers), K, rows, D, C, W,
                                                                   bias, reinterpret_cast<bf16*>(xmean), logits);
    }
    return a2v_check_launch("layer_mean_head_fwd");
}

extern "C" int a2v_head_bwd(int dtype, const float* dlogits, const void* xmean, const float* W, int K, int64_t rows, int D,
                            int C, void* g, float* dW, float* db, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "head_bwd: bad dtype");
    A2V_REQUIRE(dlogits && xmean && W && dW && K >= 1 && rows >= 0, "head_bwd: bad arguments");
    A2V_REQUIRE(D > 0 && D % 4 == 0 && C >= 1 && C <= HEAD_CMAX, "head_bwd: D %% 4 == 0 and 1 <= classes <= %d", HEAD_CMAX);
    if (rows == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long bx = rows < (long long)a2v_num_sms() * 4 ? rows : (long long)a2v_num_sms() * 4;
    dim3 grid((unsigned)bx, (unsigned)ceil_div(D, 1024));
#define A2V_HEAD_BWD(T_, CT_)                                                                                          \
    head_bwd_kernel<T_, CT_><<<grid, 256, 0, st>>>(dlogits, reinterpret_cast<const T_*>(xmean), W, K, rows, D, C,       \
                                                   reinterpret_cast<T_*>(g), dW, db)
    if (dtype == A2V_F32) {
        if (C <= 16) A2V_HEAD_BWD(float, 16); else A2V_HEAD_BWD(float, 32);
    } else {
        if (C <= 16) A2V_HEAD_BWD(bf16, 16); else A2V_HEAD_BWD(bf16, 32);
    }
#undef A2V_HEAD_BWD
    return a2v_check_launch("head_bwd");
}

static int focal_args(const float* logits, const float* targets, const int32_t* perm, int64_t rows, int C,
                      int rows_per_clip, float r, float alpha, float gamma, float threshold, FocalParams& p) {
    A2V_REQUIRE(logits && targets && rows >= 0 && C >= 1, "focal_loss: bad arguments");
    A2V_REQUIRE(perm == nullptr || (rows_per_clip > 0 && rows % rows_per_clip == 0),
                "focal_loss: target mixup needs rows_per_clip dividing rows");
    p.logits = logits; p.targets = targets; p.perm = perm; p.rows = rows; p.C = C;
    p.rows_per_clip = rows_per_clip > 0 ? rows_per_clip : 1;
    p.r = r; p.alpha = alpha; p.gamma = gamma; p.threshold = threshold;
    return A2V_OK;
}

extern "C" int a2v_focal_loss_fwd(const float* logits, const float* targets, const int32_t* perm, int64_t rows, int C,
                                  int rows_per_clip, float r, float alpha, float gamma, float threshold, double* loss_sum,
                                  float* loss_out, float* mixed_targets, uint64_t* counters, a2v_stream_t stream) {
    FocalParams p;
    int rc = focal_args(logits, targets, perm, rows, C, rows_per_clip, r, alpha, gamma, threshold, p);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(loss_sum != nullptr, "focal_loss_fwd: loss_sum is required");
    if (rows == 0) return A2V_OK;
    focal_loss_fwd_kernel<<<flat_blocks(rows * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, loss_sum, loss_out, mixed_targets, reinterpret_cast<unsigned long long*>(counters));
    return a2v_check_launch("focal_loss_fwd");
}

extern "C" int a2v_focal_loss_bwd(const float* logits, const float* targets, const int32_t* perm, int64_t rows, int C,
                                  int rows_per_clip, float r, float alpha, float gamma, const float* grad_out,
                                  float* dlogits, a2v_stream_t stream) {
    FocalParams p;
    int rc = focal_args(logits, targets, perm, rows, C, rows_per_clip, r, alpha, gamma, 0.f, p);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(dlogits != nullptr, "focal_loss_bwd: dlogits is required");
    if (rows == 0) return A2V_OK;
    focal_loss_bwd_kernel<<<flat_blocks(rows * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, grad_out,
                                                                                                          dlogits);
    return a2v_check_launch("focal_loss_bwd");
}

extern "C" int a2v_channel_mask(int dtype, void* x, const uint8_t* chmask, int64_t rows, int rows_per_clip, int D,
                                a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "channel_mask: bad dtype");
    A2V_REQUIRE(x && chmask && rows >= 0 && rows_per_clip > 0 && D > 0 && D % 4 == 0, "channel_mask: bad arguments");
    if (rows == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_blocks(rows * (D / 4), 256);
    if (dtype == A2V_F32)
        channel_mask_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(x), chmask, rows, rows_per_clip, D);
    else
        channel_mask_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<bf16*>(x), chmask, rows, rows_per_clip, D);
    return a2v_check_launch("channel_mask");
}
