// Shared pieces of the attention kernels (attention.cu: forward + resident backward, attention_bwd.cu: tiled
// backward for any length): parameter block, ALiBi coefficient, the dropout hash every kernel derives its keep flags
// from, descriptor validation.
#pragma once
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int HD = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
    const void* qkv;
    void* out;
    float* lse;
    const int* pos;
    const float* slopes;
    const float* alibi_scale;
    int alibi_scale_stride;
    int batch, L, H, D;
    float sm_scale;
    float drop_p;
    unsigned long long seed;
    const void* dout;
    void* dqkv;
    float* dalibi_scale;
    const float* qk_bound;  // optional, (batch * H) x {max |q|^2, max |k|^2} of the head (see attn_qk_bound_kernel)
    // tiled backward (attention_bwd.cu): workspace written by a2v_attn_bwd_prepare
    const float* delta;     // (batch, H, L) rowsum(dO * O)
    float* dq_acc;          // (batch, L, D) fp32 accumulator of dQ (unscaled)
};

__device__ __forceinline__ float head_coef(const AttnParams& p, int h) {
    float sc = 1.0f;
    if (p.alibi_scale != nullptr) sc = fmaxf(p.alibi_scale[h * p.alibi_scale_stride], 0.f);
    return p.slopes != nullptr ? p.slopes[h] * sc : 0.f;
}

// Attention-dropout bits: 16 bits per (query row, key), generated four keys at a time from 32-bit
// multiply-xorshift hashes of a per-row key (one hash pair per group of 4 keys, ~3.5 instructions per
// probability). Every kernel of this file (forward, both backwards, fp32 validation) derives its keep
// flags from these two functions, so forward and backward always agree.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t attn_row_key(unsigned long long seed, long long bh, int L, int i) {
    return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + (uint32_t)(bh * L + i)));
}
// .x: keys 4g, 4g+1 (low / high half), .y: keys 4g+2, 4g+3
__device__ __forceinline__ uint2 attn_bits4(uint32_t row_key, int g) {
    uint32_t a = row_key + (uint32_t)g * 0x9E3779B9U;
    uint32_t b = a + 0x85ebca6bU;
    a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16;
    b *= 0x7feb352dU; b ^= b >> 15; b *= 0x846ca68bU; b ^= b >> 16;
    return make_uint2(a, b);
}
__device__ __forceinline__ uint32_t attn_drop_threshold(float pd) { return (uint32_t)(pd * 65536.0f); }
__device__ __forceinline__ bool attn_keep(unsigned long long seed, long long bh, int L, int i, int j, float pd) {
    const uint2 bits = attn_bits4(attn_row_key(seed, bh, L, i), j >> 2);
    const uint32_t w = (j & 2) ? bits.y : bits.x;
    return ((w >> (16 * (j & 1))) & 0xffffu) >= attn_drop_threshold(pd);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); NULL when unavailable
EncodeTiledFn2 attn_tensor_map_encoder();
// 3-D bf16 map over (cols, L, batch) with a {64, box_rows, 1} box, 128-byte swizzle
int attn_make_map(CUtensorMap* out, const void* base, int cols, int L, int batch, int box_rows);
int validate_attn(const a2v_attn_desc* d, AttnParams& p);
// attention_bwd.cu: tiled backward for any length (p.delta / p.dq_acc set from the caller's workspace)
int attn_bwd_tiled_launch(const AttnParams& p, cudaStream_t st);

}  // namespace a2v
