// Shared device helpers for the sm_100a kernels: PTX wrappers for mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), warp reductions,
// a counter-based RNG for dropout / mask-token noise, and the C-ABI error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define A2V_OK 0
#define A2V_ERR_ARG 1
#define A2V_ERR_CUDA 2
#define A2V_ERR_UNSUPPORTED 3

void a2v_set_error(const char* fmt, ...);
int a2v_check_launch(const char* what);
// opt a kernel into `bytes` of dynamic shared memory on the CURRENT device (once per function and device, thread safe)
int a2v_ensure_dynamic_smem(const void* func, size_t bytes);

#define A2V_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            a2v_set_error(__VA_ARGS__);                \
            return A2V_ERR_ARG;                        \
        }                                              \
    } while (0)

namespace a2v {

typedef __nv_bfloat16 bf16;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------
// generic loads / stores with type conversion
// ----------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// load / store 4 consecutive elements (16 B for float, 8 B for bf16); pointers must be aligned
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
    uint2 t;
    t.x = pack_bf16x2(v[0], v[1]);
    t.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
}

// ----------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------
__device__ __forceinline__ float gelu_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_exact_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// erf via Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7): one rcp + one ex2 instead of erff's
// long polynomial. Used by the bf16 kernels (the fp32 validation kernels keep erff).
//   erf(u) = sign(u) * (1 - poly(t) * exp(-u^2)),  t = 1 / (1 + 0.3275911 |u|)
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 14-instruction GELU (erfc form): y = relu(x) - |x| * erfc(|x|/sqrt2)/2.
// h(x) = erfc(|x|/sqrt2)/2 (Abramowitz-Stegun 7.1.26, halved coefficients) and e = exp(-x^2/2)
__device__ __forceinline__ float half_erfc(float x, float& e) {
    const float a = fabsf(x);
    const float t = rcp_approx(fmaf(0.23164188f, a, 1.0f));
    e = ex2_approx_f((x * -0.72134752f) * x);
    float pl = fmaf(0.5307027145f, t, -0.7265760135f);
    pl = fmaf(pl, t, 0.7107068705f);
    pl = fmaf(pl, t, -0.142248368f);
    pl = fmaf(pl, t, 0.127414796f);
    return (pl * t) * e;
}
__device__ __forceinline__ float gelu_erfc(float x) {
    float e;
    const float h = half_erfc(x, e);
    return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erfc_grad(float x) {
    float e;
    const float h = half_erfc(x, e);
    const float cdf = x >= 0.f ? 1.0f - h : h;
    return fmaf(x * 0.3989422804014327f, e, cdf);
}

// 6-instruction GELU for the GEMM epilogues that write bf16 (tanh form, one MUFU): |error| vs the erf form
// <= 5e-4 absolute, a tenth of the bf16 rounding of the stored value (relative L2 error of the bf16 result against
// exact GELU: 1.67e-3 vs 1.65e-3 for the erf form). The saved pre-activation keeps the backward on the erf derivative.
__device__ __forceinline__ float tanh_approx_f(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
    const float x2 = x * x;
    const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx_f(u), hx);
}

// derivative of the tanh form: 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x), u = x (a + b x^2); |error| vs the erf form's
// derivative <= 1e-3, below the bf16 rounding of the gradient it multiplies
__device__ __forceinline__ float gelu_tanh_fast_grad(float x) {
    const float x2 = x * x;
    const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);
    const float t = tanh_approx_f(u);
    const float up = fmaf(0.1070322243f, x2, 0.7978845608f);
    const float w = (0.5f * x) * fmaf(-t, t, 1.0f);
    return fmaf(w, up, fmaf(0.5f, t, 0.5f));
}

// every bf16 kernel: the row-wise and elementwise GELU kernels are issue-bound, not HBM-bound, with the
// 14 / 17-instruction erfc forms (measured: LN+GELU over 576000 x 1024 rows 0.53 -> 0.46 ms, its backward 0.70 -> 0.62)
__device__ __forceinline__ float gelu_fast(float x) { return gelu_tanh_fast(x); }
__device__ __forceinline__ float gelu_fast_grad(float x) { return gelu_tanh_fast_grad(x); }
template <typename T> __device__ __forceinline__ float gelu_t(float x);
template <> __device__ __forceinline__ float gelu_t<float>(float x) { return gelu_exact(x); }
template <> __device__ __forceinline__ float gelu_t<bf16>(float x) { return gelu_fast(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x);
template <> __device__ __forceinline__ float gelu_grad_t<float>(float x) { return gelu_exact_grad(x); }
template <> __device__ __forceinline__ float gelu_grad_t<bf16>(float x) { return gelu_fast_grad(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ----------------------------------------------------------------------------
// counter-based RNG (splitmix/murmur style finaliser): one 64-bit hash per
// (seed, stream, index) gives two 32-bit uniforms. Used for dropout masks that
// backward regenerates instead of storing, and for the decoder mask-token noise.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rng64(uint64_t seed, uint64_t idx) {
    return mix64(seed + 0x9e3779b97f4a7c15ULL * (idx + 1));
}
// uniform in [0,1)
__device__ __forceinline__ float u01(uint32_t bits) { return (bits >> 8) * (1.0f / 16777216.0f); }
// keep-flag for dropout with drop probability p, element index idx
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t idx, float p) {
    return u01((uint32_t)rng64(seed, idx)) >= p;
}
// 4 keep flags from a single hash (16 bits each); idx4 = index of the 4-element group
__device__ __forceinline__ void drop_keep4(uint64_t seed, uint64_t idx4, float p, bool (&k)[4]) {
    const uint64_t h = rng64(seed, idx4);
    const uint32_t thr = (uint32_t)(p * 65536.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = ((uint32_t)(h >> (16 * i)) & 0xffffu) >= thr;
}

// Column sums over the 32 lanes of a warp for N columns held per lane (N = 16 or 32): a transposing butterfly, N - 1
// (+1 for N = 16) shuffles instead of 5 N. Afterwards v[0] of lane l is the sum of column l % N over all 32 lanes.
template <int N>
__device__ __forceinline__ float warp_colsum(float (&v)[N], int lane) {
#pragma unroll
    for (int s = N / 2; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = upper ? v[k] : v[k + s];
            const float keep = upper ? v[k + s] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    if (N == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
    return v[0];
}

// ----------------------------------------------------------------------------
// PTX: shared-memory address, mbarrier, fences
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// wait with a hardware suspend-time hint: the warp sleeps inside try_wait (up to the hint, woken by the phase
// completion) instead of re-issuing a poll / branch / yield triple -- a spinning warp otherwise takes issue slots from
// the working warps on its scheduler (ncu on the attention kernels: 35-40 % of all issued instructions were polls)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------
// PTX: TMA tile loads (global -> shared, completes on an mbarrier)
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ----------------------------------------------------------------------------
// PTX: tcgen05 (TMEM alloc, MMA, commit, load)
// ----------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same product with the A operand read from TENSOR MEMORY (M lanes x K/2 packed-bf16 columns): no shared-memory
// traffic for A
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, 128-byte swizzle (layout_type 2, version 1).
//  K-major : rows of 64 bf16 (128 B); 8-row groups SBO bytes apart; LBO unused.
//  MN-major: rows (along K) of 64 contiguous MN elements; 8-row groups SBO bytes apart,
//            consecutive 64-element MN chunks LBO bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// UMMA shared-memory descriptor WITHOUT swizzle (layout_type 0), K-major: 8 x 8-element core matrices of 128 contiguous
// bytes (8 rows of 16 B); the core matrix next in K is lbo_bytes away, the next 8-row group sbo_bytes away.
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}
// instruction descriptor for kind::f16 with bf16 A/B and fp32 D
__host__ __device__ inline uint32_t umma_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                        // D format fp32
    d |= 1u << 7;                        // A format bf16
    d |= 1u << 10;                       // B format bf16
    d |= (a_mn_major ? 1u : 0u) << 15;   // A major
    d |= (b_mn_major ? 1u : 0u) << 16;   // B major
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

}  // namespace a2v
