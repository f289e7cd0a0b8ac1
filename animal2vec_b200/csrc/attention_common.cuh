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
    float* dbias;           // optional (resident backward): += column sums of dqkv, (3 * D) fp32 = the qkv bias gradient
    int head_filter;        // flash forward: 1 = only the heads attention_stream.cu declined (attn_stream_head_ok false)
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

// attention_stream.cu exponentiates against a fixed per-row reference, which needs 2 max|q| max|k| scale <= 96 log2
// units; the same test (same arithmetic) lets the flash kernel pick up exactly the heads the stream kernel declined
__device__ __forceinline__ bool attn_stream_head_ok(const AttnParams& p, int b, int h) {
    const float2 b2 = reinterpret_cast<const float2*>(p.qk_bound)[b * p.H + h];  // max|q|^2, max|k|^2
    return 2.f * (sqrtf(b2.x) * sqrtf(b2.y) * 1.002f) * (p.sm_scale * LOG2E) <= 96.f;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// registers -> tensor memory (the probabilities P become the TMEM A operand of P.V)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                             uint32_t r4, uint32_t r5, uint32_t r6, uint32_t r7) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
// attention_short.cu: persistent single-pass forward for the student's short sequences (L <= 160)
int attn_fwd_short_launch(const AttnParams& p, cudaStream_t st);
constexpr int ATTN_SHORT_LMAX = 160;
// attention_stream.cu: single-pass forward for contiguous sequences with the q/k bound (p.qk_bound); heads it declines
// are left untouched for the flash kernel launched with head_filter = 1
int attn_fwd_stream_launch(const AttnParams& p, cudaStream_t st);

}  // namespace a2v
