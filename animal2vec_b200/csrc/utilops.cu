// Small HBM-bound utilities: column sums (bias gradients), strided cast/permute (weight
// layouts for the GEMM), bf16 hi/lo splitting (validation "fp32 mode" of the GEMM),
// fused EMA teacher update, fused AdamW, gradient norm / clip coefficient.
//
// Reference behaviour replaced:
//   fairseq EMAModule.step + load_state_dict (called nn/data2vec2.py:408)  -> a2v_ema_step
//   fairseq Adam (decoupled weight decay) + clip_grad_norm (yaml clip_norm) -> a2v_adamw_step,
//   a2v_sumsq, a2v_clip_coef
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

// ---------------------------------------------------------------- column sums
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long rows,
                                                     int C, long long rows_per_block) {
    __shared__ float part[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + lane * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        for (long long r = r0 + warp; r < r1; r += 8) {
            float v[4];
            load4(x + r * C + c, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[warp][lane * 4 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
        const int cc = blockIdx.x * 128 + threadIdx.x;
        if (cc < C) atomicAdd(out + cc, s);
    }
}

// bf16, C % 256 == 0 (every bias-gradient site of the step): 16-byte loads, four independent rows in flight
// per warp so the ~64 KB per SM that HBM latency needs are outstanding
__global__ void __launch_bounds__(256) colsum_bf16_wide_kernel(const bf16* __restrict__ x, float* __restrict__ out,
                                                               long long rows, int C, long long rows_per_block) {
    __shared__ float part[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 256 + lane * 8;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bf16* base = x + c;
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (r + 8 * u) * C));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 a = unpack_bf16x2(v[u].x), b = unpack_bf16x2(v[u].y), d = unpack_bf16x2(v[u].z),
                         e = unpack_bf16x2(v[u].w);
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
            acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
        }
    }
    for (; r < r1; r += 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + r * C));
        const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), d = unpack_bf16x2(v.z), e = unpack_bf16x2(v.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
    atomicAdd(out + blockIdx.x * 256 + threadIdx.x, s);
}

// ---------------------------------------------------------------- strided cast / permute (4-D)
struct PermuteParams {
    const void* in;
    void* out;
    long long dims[4];
    long long in_strides[4];
    long long in_offset;
    long long total;
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_strided_kernel(const PermuteParams p) {
    const TI* in = reinterpret_cast<const TI*>(p.in);
    TO* out = reinterpret_cast<TO*>(p.out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        long long rem = i;
        const long long i3 = rem % p.dims[3]; rem /= p.dims[3];
        const long long i2 = rem % p.dims[2]; rem /= p.dims[2];
        const long long i1 = rem % p.dims[1]; rem /= p.dims[1];
        const long long i0 = rem;
        const long long src = p.in_offset + i0 * p.in_strides[0] + i1 * p.in_strides[1] + i2 * p.in_strides[2] +
                              i3 * p.in_strides[3];
        out[i] = from_f32<TO>(src >= 0 ? to_f32(in[src]) : 0.f);
    }
}

// ---------------------------------------------------------------- general 4-D re-layout
// out[out_offset + sum i_d*out_strides[d]] (+)= in[in_offset + sum i_d*in_strides[d]]
// (weight packing into padded / tap-major / transposed GEMM layouts and the inverse for
// the weight gradients; destinations with pad slots are zero-initialised once by the caller)
struct RelayoutParams {
    const void* in;
    void* out;
    long long dims[4];
    long long in_strides[4];
    long long out_strides[4];
    long long in_offset, out_offset;
    long long total;
    int accumulate;
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) relayout_kernel(const RelayoutParams p) {
    const TI* in = reinterpret_cast<const TI*>(p.in);
    TO* out = reinterpret_cast<TO*>(p.out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        long long rem = i;
        const long long i3 = rem % p.dims[3]; rem /= p.dims[3];
        const long long i2 = rem % p.dims[2]; rem /= p.dims[2];
        const long long i1 = rem % p.dims[1]; rem /= p.dims[1];
        const long long i0 = rem;
        const long long src = p.in_offset + i0 * p.in_strides[0] + i1 * p.in_strides[1] + i2 * p.in_strides[2] +
                              i3 * p.in_strides[3];
        const long long dst = p.out_offset + i0 * p.out_strides[0] + i1 * p.out_strides[1] + i2 * p.out_strides[2] +
                              i3 * p.out_strides[3];
        float v = to_f32(in[src]);
        if (p.accumulate) v += to_f32(out[dst]);
        out[dst] = from_f32<TO>(v);
    }
}

// Batched re-layout: one launch walks a DEVICE-resident table of re-layout items (blockIdx.y = item). The
// weight packs of a model are rebuilt after every optimizer / EMA update (about 150 items per step at the
// large config, most of them a few MB); one table-driven launch replaces as many dependent small launches.
// 32-bit index arithmetic (every item is far below 2^31 elements; checked when the table is built).
template <typename TI, typename TO>
__device__ __forceinline__ void relayout_item_body(const a2v_relayout_item& it) {
    const TI* in = reinterpret_cast<const TI*>(it.in);
    TO* out = reinterpret_cast<TO*>(it.out);
    const uint32_t d1 = (uint32_t)it.dims[1], d2 = (uint32_t)it.dims[2], d3 = (uint32_t)it.dims[3];
    const uint32_t total = (uint32_t)(it.dims[0] * it.dims[1] * it.dims[2] * it.dims[3]);
    const long long is0 = it.in_strides[0], is1 = it.in_strides[1], is2 = it.in_strides[2], is3 = it.in_strides[3];
    const long long os0 = it.out_strides[0], os1 = it.out_strides[1], os2 = it.out_strides[2], os3 = it.out_strides[3];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t rem = i;
        const uint32_t i3 = rem % d3; rem /= d3;
        const uint32_t i2 = rem % d2; rem /= d2;
        const uint32_t i1 = rem % d1; rem /= d1;
        const long long src = it.in_offset + rem * is0 + i1 * is1 + i2 * is2 + i3 * is3;
        const long long dst = it.out_offset + rem * os0 + i1 * os1 + i2 * os2 + i3 * os3;
        float v = to_f32(in[src]);
        if (it.accumulate) v += to_f32(out[dst]);
        out[dst] = from_f32<TO>(v);
        if (it.zero_src) const_cast<TI*>(in)[src] = from_f32<TI>(0.f);
    }
}

// 2-D transposing items (dims (1, 1, K, N), input contiguous along K, output contiguous along N: the W -> W^T packs
// of every Linear): 32x32 tiles through shared memory so that both the reads and the writes are coalesced.
template <typename TI, typename TO>
__device__ __forceinline__ void relayout_item_transpose(const a2v_relayout_item& it, float (*tile)[33]) {
    const TI* in = reinterpret_cast<const TI*>(it.in) + it.in_offset;
    TO* out = reinterpret_cast<TO*>(it.out) + it.out_offset;
    const int K = (int)it.dims[2], N = (int)it.dims[3];
    const long long is3 = it.in_strides[3], os2 = it.out_strides[2];
    const int tk = (K + 31) / 32, tn = (N + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int t = blockIdx.x; t < tk * tn; t += gridDim.x) {
        const int k0 = (t % tk) * 32, n0 = (t / tk) * 32;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = n0 + ty + 8 * r, k = k0 + tx;
            tile[ty + 8 * r][tx] = (n < N && k < K) ? to_f32(in[n * is3 + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = k0 + ty + 8 * r, n = n0 + tx;
            if (k < K && n < N) {
                float v = tile[tx][ty + 8 * r];
                if (it.accumulate) v += to_f32(out[k * os2 + n]);
                out[k * os2 + n] = from_f32<TO>(v);
            }
        }
        __syncthreads();
    }
}

// The same for fp32 -> bf16 without accumulation (the transposed weight copies of every Linear, 300 M elements per
// update): 64 x 64 tiles, 16-byte loads along K, 4-byte (bf16 pair) stores along N -- four times the bytes in flight per
// block of the 32 x 32 version, which ran at 1.2 TB/s (ncu, profiles/r2_ncu_summary.md).
__device__ __forceinline__ void relayout_item_transpose64(const a2v_relayout_item& it, float (*tile)[65]) {
    const float* in = reinterpret_cast<const float*>(it.in) + it.in_offset;
    bf16* out = reinterpret_cast<bf16*>(it.out) + it.out_offset;
    const int K = (int)it.dims[2], N = (int)it.dims[3];
    const long long is3 = it.in_strides[3], os2 = it.out_strides[2];
    const int tk = K / 64, tn = N / 64;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;   // load: 16 threads x float4 per row, 16 rows per pass
    const int sx = threadIdx.x & 31, sy = threadIdx.x >> 5;   // store: 32 threads x bf16 pair per row, 8 rows per pass
    for (int t = blockIdx.x; t < tk * tn; t += gridDim.x) {
        const int k0 = (t % tk) * 64, n0 = (t / tk) * 64;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = ly + 16 * r;
            const float4 v = *reinterpret_cast<const float4*>(in + (long long)(n0 + n) * is3 + k0 + 4 * lx);
            tile[n][4 * lx + 0] = v.x;
            tile[n][4 * lx + 1] = v.y;
            tile[n][4 * lx + 2] = v.z;
            tile[n][4 * lx + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int k = sy + 8 * r;
            const uint32_t pk = pack_bf16x2(tile[2 * sx][k], tile[2 * sx + 1][k]);
            *reinterpret_cast<uint32_t*>(out + (long long)(k0 + k) * os2 + n0 + 2 * sx) = pk;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) relayout_batch_kernel(const a2v_relayout_item* __restrict__ items) {
    __shared__ a2v_relayout_item it;
    __shared__ __align__(16) float tile_raw[64 * 65];
    float (*tile)[33] = reinterpret_cast<float (*)[33]>(tile_raw);
    if (threadIdx.x == 0) it = items[blockIdx.y];
    __syncthreads();
    if (it.dims[0] == 1 && it.dims[1] == 1 && it.in_strides[2] == 1 && it.out_strides[3] == 1 && !it.zero_src &&
        !it.accumulate && it.in_dtype == A2V_F32 && it.out_dtype == A2V_BF16 && it.dims[2] % 64 == 0 &&
        it.dims[3] % 64 == 0 && it.in_strides[3] % 4 == 0 && it.in_offset % 4 == 0 && it.out_strides[2] % 2 == 0 &&
        it.out_offset % 2 == 0 && (reinterpret_cast<uintptr_t>(it.in) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(it.out) & 3) == 0) {
        relayout_item_transpose64(it, reinterpret_cast<float (*)[65]>(tile_raw));
        return;
    }
    if (it.dims[0] == 1 && it.dims[1] == 1 && it.in_strides[2] == 1 && it.out_strides[3] == 1 && !it.zero_src &&
        it.dims[2] >= 32 && it.dims[3] >= 32) {
        if (it.in_dtype == A2V_F32 && it.out_dtype == A2V_BF16) { relayout_item_transpose<float, bf16>(it, tile); return; }
        if (it.in_dtype == A2V_F32 && it.out_dtype == A2V_F32) { relayout_item_transpose<float, float>(it, tile); return; }
    }
    if (it.in_dtype == A2V_F32 && it.out_dtype == A2V_BF16) relayout_item_body<float, bf16>(it);
    else if (it.in_dtype == A2V_F32 && it.out_dtype == A2V_F32) relayout_item_body<float, float>(it);
    else if (it.in_dtype == A2V_BF16 && it.out_dtype == A2V_BF16) relayout_item_body<bf16, bf16>(it);
    else if (it.in_dtype == A2V_BF16 && it.out_dtype == A2V_F32) relayout_item_body<bf16, float>(it);
}

// ---------------------------------------------------------------- flat elementwise
// out = dh * GELU'(u): backward of timm Mlp's activation (modules.py:312-317), applied to the output of
// the fc2 data-gradient GEMM (kept out of that GEMM's epilogue: a latency-bound read there costs 5x more).
template <typename T>
__global__ void __launch_bounds__(256) dgelu_mul_kernel(const T* __restrict__ dh, const T* __restrict__ u,
                                                        T* __restrict__ out, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        load4(dh + 8 * i, *reinterpret_cast<float(*)[4]>(a));
        load4(dh + 8 * i + 4, *reinterpret_cast<float(*)[4]>(a + 4));
        load4(u + 8 * i, *reinterpret_cast<float(*)[4]>(b));
        load4(u + 8 * i + 4, *reinterpret_cast<float(*)[4]>(b + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= gelu_grad_t<T>(b[j]);
        store4(out + 8 * i, *reinterpret_cast<float(*)[4]>(a));
        store4(out + 8 * i + 4, *reinterpret_cast<float(*)[4]>(a + 4));
    }
}

// The same with the column sums of the result fused in (the bias gradient of fc1, rows x C with C % 8 == 0): the
// launch uses a thread count that is a multiple of C / 8, so a thread keeps ONE group of 8 columns over its
// whole grid-stride loop and accumulates their sums in registers; one 16-byte vector reduction pair per thread at
// the end. Saves a separate pass over the (rows x 4 D) tensor.
template <typename T>
__global__ void __launch_bounds__(256) dgelu_mul_colsum_kernel(const T* __restrict__ dh, const T* __restrict__ u,
                                                               T* __restrict__ out, long long n8, int cols8,
                                                               float* __restrict__ colsum) {
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long i = t0; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        load4(dh + 8 * i, *reinterpret_cast<float(*)[4]>(a));
        load4(dh + 8 * i + 4, *reinterpret_cast<float(*)[4]>(a + 4));
        load4(u + 8 * i, *reinterpret_cast<float(*)[4]>(b));
        load4(u + 8 * i + 4, *reinterpret_cast<float(*)[4]>(b + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] *= gelu_grad_t<T>(b[j]);
            a[j] = to_f32(from_f32<T>(a[j]));  // sum what is stored
            acc[j] += a[j];
        }
        store4(out + 8 * i, *reinterpret_cast<float(*)[4]>(a));
        store4(out + 8 * i + 4, *reinterpret_cast<float(*)[4]>(a + 4));
    }
    float* dst = colsum + (t0 % cols8) * 8;
    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(acc[0], acc[1], acc[2], acc[3]));
    atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                            long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float v[4];
        load4(in + 4 * i, v);
        store4(out + 4 * i, v);
    }
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); written K-concatenated per row as
// pattern 0: [hi | hi | lo], pattern 1: [hi | lo | hi] so that an ordinary bf16 GEMM over 3K
// computes hi*hi + hi*lo + lo*hi (about 16 mantissa bits) -- the validation "fp32 mode".
// pattern 2/3: the same stacked along rows (for MN-major / reduction-over-rows products).
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                     long long rows, int K, int pattern) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * K;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K;
        const int k = (int)(i - r * K);
        const float x = in[i];
        const bf16 hi = __float2bfloat16_rn(x);
        const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        const bf16 second = (pattern & 1) ? lo : hi;
        const bf16 third = (pattern & 1) ? hi : lo;
        if (pattern < 2) {
            bf16* o = out + r * 3 * K;
            o[k] = hi;
            o[K + k] = second;
            o[2 * K + k] = third;
        } else {
            out[i] = hi;
            out[rows * K + i] = second;
            out[2 * rows * K + i] = third;
        }
    }
}

// ---------------------------------------------------------------- EMA teacher update
// shadow = decay * shadow + (1 - decay) * student   (fp32 master of the teacher)
// teacher_lp = bf16(shadow)                         (the copy the teacher GEMMs read)
// 4 B read (student) + 4 B read + 4 B write (shadow) + 2 B write (bf16) per parameter.
__global__ void __launch_bounds__(256) ema_kernel(const float* __restrict__ student, float* __restrict__ shadow,
                                                  bf16* __restrict__ teacher_lp, long long n4, float decay) {
    const float om = 1.0f - decay;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float s[4], e[4];
        load4(student + 4 * i, s);
        load4(shadow + 4 * i, e);
#pragma unroll
        for (int j = 0; j < 4; ++j) e[j] = e[j] * decay + s[j] * om;  // mul_(decay).add_(p, alpha=1-decay)
        store4(shadow + 4 * i, e);
        if (teacher_lp != nullptr) store4(teacher_lp + 4 * i, e);
    }
}

// ---------------------------------------------------------------- AdamW (fairseq Adam semantics)
struct AdamParams {
    float* p;
    const float* g;
    float* m;
    float* v;
    bf16* p_lp;
    long long n4;
    float lr, beta1, beta2, eps, wd, step_size;
    const float* grad_scale;  // device scalar multiplied into every gradient (clip * 1/sample_size), may be NULL
    const uint8_t* wd_mask;   // per 4-element chunk: 0 = the weight_decay_scale-0 group (data2vec2.py:318-322), may be NULL
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
    const float gs = a.grad_scale != nullptr ? *a.grad_scale : 1.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
        float p[4], g[4], m[4], v[4];
        load4(a.p + 4 * i, p);
        load4(a.g + 4 * i, g);
        load4(a.m + 4 * i, m);
        load4(a.v + 4 * i, v);
        const float wd = (a.wd_mask == nullptr || a.wd_mask[i] != 0) ? a.wd : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = g[j] * gs;
            m[j] = a.beta1 * m[j] + (1.f - a.beta1) * gj;
            v[j] = a.beta2 * v[j] + (1.f - a.beta2) * gj * gj;
            const float denom = sqrtf(v[j]) + a.eps;
            p[j] = p[j] - a.lr * wd * p[j];      // decoupled weight decay: p.add_(p, alpha=-wd*lr)
            p[j] = p[j] - a.step_size * m[j] / denom;
        }
        store4(a.p + 4 * i, p);
        store4(a.m + 4 * i, m);
        store4(a.v + 4 * i, v);
        if (a.p_lp != nullptr) store4(a.p_lp + 4 * i, p);
    }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
    __shared__ float wsum[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        atomicAdd(out, (double)s);
    }
}

// out[0] = grad_mult * min(1, max_norm / (grad_mult * sqrt(sumsq) + 1e-6)), out[1] = grad_mult * sqrt(sumsq)
// grad_mult = numer / denom[0] (e.g. world_size / sample_size) when denom is given.
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, const float* __restrict__ denom, float numer,
                                 float max_norm, float* __restrict__ out) {
    float mult = numer;
    if (denom != nullptr) mult = numer / fmaxf(*denom, 1e-20f);
    const float norm = mult * (float)sqrt(*sumsq);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    out[0] = mult * coef;
    out[1] = norm;
}

// out[8 i .. 8 i + 7] = float(in[...]): the bf16 gradient buckets coming back from the all-reduce
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out,
                                                            long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = *reinterpret_cast<const uint4*>(in + 8 * i);
        const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
        *reinterpret_cast<float4*>(out + 8 * i) = make_float4(a.x, a.y, b.x, b.y);
        *reinterpret_cast<float4*>(out + 8 * i + 4) = make_float4(c.x, c.y, d.x, d.y);
    }
}

static int flat_grid(long long n) {
    long long b = ceil_div64(n, 256);
    long long cap = (long long)a2v_num_sms() * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_colsum(int dtype, const void* x, float* out, int64_t rows, int C, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "colsum: bad dtype");
    A2V_REQUIRE(x && out && C > 0 && C % 4 == 0 && rows >= 0, "colsum: bad arguments");
    if (rows == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_BF16 && C % 256 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int cb = C / 256;
        int rb = (int)((long long)a2v_num_sms() * 4 / cb);
        if (rb < 1) rb = 1;
        if (rb > rows / 32 + 1) rb = (int)(rows / 32 + 1);
        const long long rpb = ceil_div64(rows, rb);
        colsum_bf16_wide_kernel<<<dim3(cb, rb), 256, 0, st>>>((const bf16*)x, out, rows, C, rpb);
        return a2v_check_launch("colsum");
    }
    const int cblocks = ceil_div(C, 128);
    int rblocks = (int)((long long)a2v_num_sms() * 4 / cblocks);
    if (rblocks < 1) rblocks = 1;
    if (rblocks > rows / 8 + 1) rblocks = (int)(rows / 8 + 1);
    const long long rpb = ceil_div64(rows, rblocks);
    dim3 grid(cblocks, rblocks);
    if (dtype == A2V_F32)
        colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, rows, C, rpb);
    else
        colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, out, rows, C, rpb);
    return a2v_check_launch("colsum");
}

extern "C" int a2v_cast_strided(int in_dtype, int out_dtype, const void* in, void* out, const int64_t* dims4,
                                const int64_t* in_strides4, int64_t in_offset, a2v_stream_t stream) {
    A2V_REQUIRE(in && out && dims4 && in_strides4, "cast_strided: NULL pointer");
    PermuteParams p;
    p.in = in;
    p.out = out;
    p.total = 1;
    for (int i = 0; i < 4; ++i) {
        A2V_REQUIRE(dims4[i] > 0, "cast_strided: dims must be positive");
        p.dims[i] = dims4[i];
        p.in_strides[i] = in_strides4[i];
        p.total *= dims4[i];
    }
    p.in_offset = in_offset;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_grid(p.total);
    if (in_dtype == A2V_F32 && out_dtype == A2V_BF16)
        cast_strided_kernel<float, bf16><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_F32 && out_dtype == A2V_F32)
        cast_strided_kernel<float, float><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_BF16 && out_dtype == A2V_BF16)
        cast_strided_kernel<bf16, bf16><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_BF16 && out_dtype == A2V_F32)
        cast_strided_kernel<bf16, float><<<grid, 256, 0, st>>>(p);
    else {
        a2v_set_error("cast_strided: bad dtypes");
        return A2V_ERR_ARG;
    }
    return a2v_check_launch("cast_strided");
}

extern "C" int a2v_relayout_batch(const a2v_relayout_item* items_device, int n_items, int blocks_per_item,
                                  a2v_stream_t stream) {
    A2V_REQUIRE(items_device != nullptr && n_items >= 0 && n_items <= 65535, "relayout_batch: bad table (n=%d)", n_items);
    A2V_REQUIRE(blocks_per_item >= 1 && blocks_per_item <= 4096, "relayout_batch: blocks_per_item out of range");
    if (n_items == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    relayout_batch_kernel<<<dim3(blocks_per_item, n_items), 256, 0, st>>>(items_device);
    return a2v_check_launch("relayout_batch");
}

extern "C" int a2v_relayout(int in_dtype, int out_dtype, const void* in, void* out, const int64_t* dims4,
                            const int64_t* in_strides4, int64_t in_offset, const int64_t* out_strides4,
                            int64_t out_offset, int accumulate, a2v_stream_t stream) {
    A2V_REQUIRE(in && out && dims4 && in_strides4 && out_strides4, "relayout: NULL pointer");
    RelayoutParams p;
    p.in = in;
    p.out = out;
    p.total = 1;
    for (int i = 0; i < 4; ++i) {
        A2V_REQUIRE(dims4[i] > 0, "relayout: dims must be positive");
        p.dims[i] = dims4[i];
        p.in_strides[i] = in_strides4[i];
        p.out_strides[i] = out_strides4[i];
        p.total *= dims4[i];
    }
    p.in_offset = in_offset;
    p.out_offset = out_offset;
    p.accumulate = accumulate;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_grid(p.total);
    if (in_dtype == A2V_F32 && out_dtype == A2V_BF16)
        relayout_kernel<float, bf16><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_F32 && out_dtype == A2V_F32)
        relayout_kernel<float, float><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_BF16 && out_dtype == A2V_BF16)
        relayout_kernel<bf16, bf16><<<grid, 256, 0, st>>>(p);
    else if (in_dtype == A2V_BF16 && out_dtype == A2V_F32)
        relayout_kernel<bf16, float><<<grid, 256, 0, st>>>(p);
    else {
        a2v_set_error("relayout: bad dtypes");
        return A2V_ERR_ARG;
    }
    return a2v_check_launch("relayout");
}

extern "C" int a2v_dgelu_mul(int dtype, const void* dh, const void* u, void* out, int64_t n, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "dgelu_mul: bad dtype");
    A2V_REQUIRE(dh && u && out && n >= 0 && n % 8 == 0, "dgelu_mul: n must be a multiple of 8");
    if (n == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_grid(n / 8);
    if (dtype == A2V_F32)
        dgelu_mul_kernel<float><<<grid, 256, 0, st>>>((const float*)dh, (const float*)u, (float*)out, n / 8);
    else
        dgelu_mul_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dh, (const bf16*)u, (bf16*)out, n / 8);
    return a2v_check_launch("dgelu_mul");
}

extern "C" int a2v_dgelu_mul_colsum(int dtype, const void* dh, const void* u, void* out, int64_t rows, int C,
                                    float* colsum, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "dgelu_mul_colsum: bad dtype");
    A2V_REQUIRE(dh && u && out && colsum && rows >= 0 && C > 0 && C % 8 == 0, "dgelu_mul_colsum: C must be a multiple of 8");
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(colsum) & 15) == 0, "dgelu_mul_colsum: colsum must be 16-byte aligned");
    if (rows == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // thread count = multiple of C / 8: grid = k * (cols8 / gcd(cols8, 256))
    const int cols8 = C / 8;
    int g = cols8, h = 256;
    while (h) { const int r = g % h; g = h; h = r; }
    const int unit = cols8 / g;
    const long long want = (long long)a2v_num_sms() * 8;
    const long long need = ceil_div64(rows * cols8, 256);
    long long grid = (want < need ? want : need) / unit * unit;
    if (grid < unit) grid = unit;
    A2V_REQUIRE(grid <= 65535 * 16, "dgelu_mul_colsum: unsupported column count %d", C);
    if (dtype == A2V_F32)
        dgelu_mul_colsum_kernel<float><<<(int)grid, 256, 0, st>>>((const float*)dh, (const float*)u, (float*)out,
                                                                 rows * cols8, cols8, colsum);
    else
        dgelu_mul_colsum_kernel<bf16><<<(int)grid, 256, 0, st>>>((const bf16*)dh, (const bf16*)u, (bf16*)out,
                                                                rows * cols8, cols8, colsum);
    return a2v_check_launch("dgelu_mul_colsum");
}

extern "C" int a2v_cast_f32_to_bf16(const float* in, void* out, int64_t n, a2v_stream_t stream) {
    A2V_REQUIRE(in && out && n >= 0 && n % 4 == 0, "cast_f32_to_bf16: n must be a multiple of 4");
    if (n == 0) return A2V_OK;
    cast_f32_bf16_kernel<<<flat_grid(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, (bf16*)out, n / 4);
    return a2v_check_launch("cast_f32_to_bf16");
}

extern "C" int a2v_cast_bf16_to_f32(const void* in, float* out, int64_t n, a2v_stream_t stream) {
    A2V_REQUIRE(in && out && n >= 0 && n % 8 == 0, "cast_bf16_to_f32: n must be a multiple of 8");
    A2V_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "cast_bf16_to_f32: 16-byte alignment");
    if (n == 0) return A2V_OK;
    cast_bf16_f32_kernel<<<flat_grid(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)in, out, n / 8);
    return a2v_check_launch("cast_bf16_to_f32");
}

extern "C" int a2v_split3(const float* in, void* out, int64_t rows, int K, int pattern, a2v_stream_t stream) {
    A2V_REQUIRE(in && out && rows >= 0 && K > 0 && pattern >= 0 && pattern <= 3, "split3: bad arguments");
    if (rows == 0) return A2V_OK;
    split3_kernel<<<flat_grid(rows * K), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, (bf16*)out, rows, K,
                                                                                          pattern);
    return a2v_check_launch("split3");
}

extern "C" int a2v_ema_step(const float* student, float* shadow, void* teacher_bf16, int64_t n, float decay,
                            a2v_stream_t stream) {
    A2V_REQUIRE(student && shadow && n >= 0 && n % 4 == 0, "ema_step: n must be a multiple of 4");
    A2V_REQUIRE(((uintptr_t)student & 15) == 0 && ((uintptr_t)shadow & 15) == 0 && ((uintptr_t)teacher_bf16 & 7) == 0,
                "ema_step: buffers must be 16-byte aligned");
    if (n == 0) return A2V_OK;
    ema_kernel<<<flat_grid(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(student, shadow,
                                                                                     (bf16*)teacher_bf16, n / 4, decay);
    return a2v_check_launch("ema_step");
}

extern "C" int a2v_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step,
                              const float* grad_scale, const uint8_t* wd_mask, a2v_stream_t stream) {
    A2V_REQUIRE(p && g && m && v && n >= 0 && n % 4 == 0 && step >= 1, "adamw_step: bad arguments");
    if (n == 0) return A2V_OK;
    AdamParams a;
    a.p = p; a.g = g; a.m = m; a.v = v; a.p_lp = (bf16*)p_bf16; a.n4 = n / 4;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    a.step_size = (float)((double)lr * sqrt(bc2) / bc1);
    a.grad_scale = grad_scale;
    a.wd_mask = wd_mask;
    adamw_kernel<<<flat_grid(a.n4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    return a2v_check_launch("adamw_step");
}

extern "C" int a2v_sumsq(const float* x, int64_t n, double* out, a2v_stream_t stream) {
    A2V_REQUIRE(x && out && n >= 0, "sumsq: bad arguments");
    if (n == 0) return A2V_OK;
    sumsq_kernel<<<flat_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out);
    return a2v_check_launch("sumsq");
}

extern "C" int a2v_clip_coef(const double* sumsq, const float* denom, float numer, float max_norm, float* out2,
                             a2v_stream_t stream) {
    A2V_REQUIRE(sumsq && out2, "clip_coef: NULL pointer");
    clip_coef_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sumsq, denom, numer, max_norm, out2);
    return a2v_check_launch("clip_coef");
}
