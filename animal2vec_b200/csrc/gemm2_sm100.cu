// CTA-pair (cta_group::2) variant of the NT GEMM for the plain Linear shapes of the transformer blocks:
//   C[m, n] = alpha * sum_k A[m, k] * W[n, k]  (+ bias[n]) -> (preact) -> GELU -> (+ residual),  bf16 in / bf16 out.
// Replaces the same nn.Linear dispatches as gemm_sm100.cu (reference nn/modalities/modules.py:368-410 attn.qkv /
// attn.proj, timm Mlp fc1 / fc2 used at modules.py:312-317, and their data gradients).
//
// Why: with one CTA per 128 x 256 tile every tcgen05.mma reads 4 KB of A and 8 KB of B per 128 tensor-pipe clocks,
// 96 B/clk of the 128 B/clk shared-memory read port; the epilogue's staging competes for the rest (ncu: its stores wait
// on the short scoreboard of the staged LDS). Two CTAs of one TPC form a pair on a 256 x 256 tile: each holds its own
// 128 rows of A and HALF of the B tile (128 of the 256 weight rows); the pair's single MMA stream reads both halves,
// so each SM streams 4 KB + 4 KB per 128 clocks (64 B/clk) and loads a third less through TMA.
//
// Roles per CTA (320 threads): warp 0 TMA producer (both CTAs load their halves; the transaction bytes of BOTH land on
// the leader's full barrier), warp 1 MMA issuer (leader CTA only; tcgen05.commit multicasts the "slot free" and
// "accumulator ready" arrivals to both CTAs), warps 2..9 epilogue out of the CTA's own 128 TMEM lanes (double-buffered
// 256-column accumulators); the peer's epilogue warps release the accumulator with a remote arrive on the leader's
// barrier.
#include <string.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int G2_BM = 128;   // rows per CTA (256 per pair)
constexpr int G2_BN = 256;   // columns per pair tile
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_THREADS = 320;
constexpr int G2_BAR_BYTES = (2 * G2_STAGES + 4) * 8 + 16;
constexpr int G2_EPI_PITCH = 80;
constexpr int G2_EPI_BYTES = 8 * 32 * G2_EPI_PITCH;
constexpr int G2_BIAS_BYTES = 8 * 128 * 4;
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + G2_BAR_BYTES + G2_EPI_BYTES + G2_BIAS_BYTES + 1024;
constexpr int G2_EPI_BIAS = 1, G2_EPI_PREACT = 2, G2_EPI_GELU = 4, G2_EPI_DGELU = 8, G2_EPI_RES = 16;
// Epilogues with GELU / GELU' arithmetic (and a second stored or loaded tile) take longer per tile than the MMAs of the
// next tile when 8 warps share them (fc1 forward 0.35 ms against 0.26 ms plain): those variants run 16 epilogue warps
// (two 32-column chunks each) and pay for the extra staging with one pipeline stage.
__host__ __device__ constexpr bool g2_heavy(int epi) { return (epi & (G2_EPI_GELU | G2_EPI_DGELU)) != 0; }
__host__ __device__ constexpr int g2_epi_warps(int epi) { return g2_heavy(epi) ? 16 : 8; }
__host__ __device__ constexpr int g2_stages(int epi) { return g2_heavy(epi) ? 5 : G2_STAGES; }
__host__ __device__ constexpr int g2_threads(int epi) { return 64 + 32 * g2_epi_warps(epi); }
__host__ __device__ constexpr int g2_smem(int epi) {
    return g2_stages(epi) * G2_STAGE_BYTES + G2_BAR_BYTES + g2_epi_warps(epi) * 32 * G2_EPI_PITCH + G2_BIAS_BYTES + 1024;
}

struct Gemm2Params {
    int M, N, kblocks;
    int n_tiles, num_tiles;
    void* c;
    long long ldc;
    float alpha;
    const float* bias;
    void* preact;
    const void* residual;  // G2_EPI_RES: added; G2_EPI_DGELU: u, the result is multiplied by GELU'(u)
    float* colsum;         // optional (G2_EPI_DGELU): += column sums of the result (the bias gradient of the Linear before)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load of a CTA pair: the bytes are written to THIS CTA's shared memory, the transaction count goes to the
// barrier at the same offset in the LEADER CTA (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
          "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs of the pair retire
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

template <int EPI>
__global__ void __launch_bounds__(g2_threads(EPI), 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Gemm2Params p) {
    constexpr int STAGES = g2_stages(EPI);
    constexpr int EW = g2_epi_warps(EPI);  // epilogue warps per CTA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * G2_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint8_t* epi_stage = smem + STAGES * G2_STAGE_BYTES + G2_BAR_BYTES;
    float* bias_stage = reinterpret_cast<float*>(epi_stage + EW * 32 * G2_EPI_PITCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * EW);  // the epilogue warps of both CTAs of the pair (used in the leader only)
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += npairs) {
                const int n_tile = tile % p.n_tiles, m2 = tile / p.n_tiles;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                    uint8_t* sb = sa + G2_A_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);  // both CTAs' halves
                    tma_load_3d_pair(sa, &tmA, &full_bar[stage], kb * G2_BK, m2 * (2 * G2_BM) + (int)rank * G2_BM, 0);
                    tma_load_3d_pair(sb, &tmB, &full_bar[stage], kb * G2_BK, n_tile * G2_BN + (int)rank * (G2_BN / 2), 0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
        if (leader && elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(2 * G2_BM, G2_BN, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += npairs) {
                mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BN;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint32_t b_base = a_base + G2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < G2_BK / 16; ++k)
                        umma2_bf16(tmem_d, umma_smem_desc(a_base + k * 32, 0, 1024), umma_smem_desc(b_base + k * 32, 0, 1024),
                                   idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma2_commit(&empty_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma2_commit(&tfull_bar[as]);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps (both CTAs), as the 8-warp epilogue
        // of gemm_sm100.cu: TMEM-native arithmetic (thread = row), bf16 staging transpose, 16-byte stores
        const int q = warp & 3;
        const int ew = warp - 2;
        constexpr int CPW = 32 / EW;  // 32-column chunks per warp: 8 warps take half of the 256 columns each, 16 a quarter
        const int c_begin = (ew >> 2) * CPW;
        const uint32_t st = smem_u32(epi_stage + ew * (32 * G2_EPI_PITCH));
        const int rrow = lane & 7, rchunk = lane >> 3;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = pair; tile < p.num_tiles; tile += npairs) {
            const int n_tile = tile % p.n_tiles, m2 = tile / p.n_tiles;
            const int row0 = m2 * (2 * G2_BM) + (int)rank * G2_BM + q * 32;
            int rows_ok = p.M - row0;
            rows_ok = rows_ok > 32 ? 32 : rows_ok;
            const long long row_off0 = (long long)row0 * p.ldc;
            const int cbase = n_tile * G2_BN + c_begin * 32;
            float* bs = bias_stage + ew * (CPW * 32);
            if (EPI & G2_EPI_BIAS) {
                if (lane < CPW * 8)
                    *reinterpret_cast<float4*>(bs + lane * 4) = __ldg(reinterpret_cast<const float4*>(p.bias + cbase) + lane);
                __syncwarp();
            }
            constexpr bool DGELU = (EPI & G2_EPI_DGELU) != 0;
            constexpr bool RES = (EPI & (G2_EPI_RES | G2_EPI_DGELU)) != 0;  // a second (M, N) bf16 operand, prefetched under the MMAs
            uint4 rsd[RES ? CPW : 1][4];
            if (RES) {
                const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + row_off0 + (long long)lane * p.ldc + cbase;
#pragma unroll
                for (int cc = 0; cc < CPW; ++cc)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rsd[RES ? cc : 0][j] = lane < rows_ok ? __ldg(reinterpret_cast<const uint4*>(rp + cc * 32) + j)
                                                              : make_uint4(0u, 0u, 0u, 0u);
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
#pragma unroll(RES ? CPW : 1)
            for (int cc = 0; cc < CPW; ++cc) {
                const int col0 = cbase + cc * 32;
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * G2_BN + (c_begin + cc) * 32, raw);
                tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bv = (EPI & G2_EPI_BIAS) ? *reinterpret_cast<const float4*>(bs + cc * 32 + 4 * j)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    x[4 * j] = fmaf(__uint_as_float(raw[4 * j]), p.alpha, bv.x);
                    x[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), p.alpha, bv.y);
                    x[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), p.alpha, bv.z);
                    x[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), p.alpha, bv.w);
                }
                if (RES) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 rv = rsd[RES ? cc : 0][j];
                        const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y);
                        const float2 r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
                        if (DGELU) {  // backward of the activation between two Linears: dh * GELU'(u)
                            x[8 * j] *= gelu_tanh_fast_grad(r0.x); x[8 * j + 1] *= gelu_tanh_fast_grad(r0.y);
                            x[8 * j + 2] *= gelu_tanh_fast_grad(r1.x); x[8 * j + 3] *= gelu_tanh_fast_grad(r1.y);
                            x[8 * j + 4] *= gelu_tanh_fast_grad(r2.x); x[8 * j + 5] *= gelu_tanh_fast_grad(r2.y);
                            x[8 * j + 6] *= gelu_tanh_fast_grad(r3.x); x[8 * j + 7] *= gelu_tanh_fast_grad(r3.y);
                        } else {
                            x[8 * j] += r0.x; x[8 * j + 1] += r0.y; x[8 * j + 2] += r1.x; x[8 * j + 3] += r1.y;
                            x[8 * j + 4] += r2.x; x[8 * j + 5] += r2.y; x[8 * j + 6] += r3.x; x[8 * j + 7] += r3.y;
                        }
                    }
                }
                if (DGELU && p.colsum != nullptr) {
                    // column sums of the result (before its bf16 rounding) over this warp's 32 rows: transposing
                    // butterfly, then one 32-lane vector of atomic adds per chunk
                    float cs[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) cs[i] = lane < rows_ok ? x[i] : 0.f;
                    const float tot = warp_colsum<32>(cs, lane);
                    atomicAdd(p.colsum + col0 + lane, tot);
                }
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 0 && !(EPI & G2_EPI_PREACT)) continue;
                    bf16* dst = reinterpret_cast<bf16*>(pass == 0 ? p.preact : p.c);
                    if (pass == 1 && (EPI & G2_EPI_GELU)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) x[i] = gelu_tanh_fast(x[i]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w0 = pack_bf16x2(x[8 * j], x[8 * j + 1]), w1 = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
                        const uint32_t w2 = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), w3 = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + lane * G2_EPI_PITCH + 16 * j),
                                     "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                                     : "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rrow + 8 * i;
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                     : "r"(st + r * G2_EPI_PITCH + 16 * rchunk)
                                     : "memory");
                        if (r < rows_ok)
                            *reinterpret_cast<uint4*>(dst + row_off0 + (long long)r * p.ldc + col0 + 8 * rchunk) = v;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TN product of a CTA pair (weight gradients of the same Linears):  out[m, n] += alpha * sum_r a[r, m] * b[r, n],
// fp32 out, split over the reduction rows with atomic accumulation. Both operands are MN-major (64-row k-blocks of
// 64-column boxes); each CTA holds 128 of the pair tile's 256 M columns and 128 of its 256 N columns. The epilogue adds
// straight out of the TMEM-native layout (thread = output row, 16-byte vector reductions): no shared-memory staging.
// ------------------------------------------------------------------------------------------------------------------
struct Gemm2TnParams {
    int M, N, kblocks, per_split;
    int n_tiles, m_tiles2, splits, num_tiles;
    float* c;
    long long ldc;
    float alpha;
};

__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2cta_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Gemm2TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + G2_STAGES;
    uint64_t* tfull_bar = empty_bar + G2_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < G2_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16);
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (n_tile, m2, split); k-block range of the split
    auto decode = [&](int tile, int& n_tile, int& m2, int& kb0, int& kb1) {
        n_tile = tile % p.n_tiles;
        tile /= p.n_tiles;
        m2 = tile % p.m_tiles2;
        const int split = tile / p.m_tiles2;
        kb0 = split * p.per_split;
        kb1 = kb0 + p.per_split;
        kb1 = kb1 < p.kblocks ? kb1 : p.kblocks;
        if (kb1 < kb0) kb1 = kb0;
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += npairs) {
                int n_tile, m2, kb0, kb1;
                decode(tile, n_tile, m2, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                    uint8_t* sb = sa + G2_A_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        tma_load_3d_pair(sa + i * 8192, &tmA, &full_bar[stage], m2 * 256 + (int)rank * 128 + i * 64, kb * G2_BK, 0);
                        tma_load_3d_pair(sb + i * 8192, &tmB, &full_bar[stage], n_tile * 256 + (int)rank * 128 + i * 64, kb * G2_BK, 0);
                    }
                    if (++stage == G2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader && elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(256, 256, true, true);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += npairs) {
                int n_tile, m2, kb0, kb1;
                decode(tile, n_tile, m2, kb0, kb1);
                mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint32_t b_base = a_base + G2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < G2_BK / 16; ++k)
                        umma2_bf16(tmem_d, umma_smem_desc(a_base + k * (16 * 128), 8192, 1024),
                                   umma_smem_desc(b_base + k * (16 * 128), 8192, 1024), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma2_commit(&empty_bar[stage]);
                    if (++stage == G2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma2_commit(&tfull_bar[as]);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int ew = warp - 2;
        const int c_begin = (ew >> 2) * 4;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = pair; tile < p.num_tiles; tile += npairs) {
            int n_tile, m2, kb0, kb1;
            decode(tile, n_tile, m2, kb0, kb1);
            const int row = m2 * 256 + (int)rank * 128 + q * 32 + lane;  // this thread's output row
            float* crow = p.c + (long long)row * p.ldc + n_tile * 256 + c_begin * 32;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            if (kb1 > kb0) {
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t raw[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * G2_BN + (c_begin + cc) * 32, raw);
                    tmem_ld_wait();
                    if (row < p.M) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            atomicAdd(reinterpret_cast<float4*>(crow + cc * 32) + j,
                                      make_float4(__uint_as_float(raw[4 * j]) * p.alpha, __uint_as_float(raw[4 * j + 1]) * p.alpha,
                                                  __uint_as_float(raw[4 * j + 2]) * p.alpha, __uint_as_float(raw[4 * j + 3]) * p.alpha));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g2_make_map(CUtensorMap* m, const a2v_operand& o, int box_rows = 128) {
    static EncodeTiledFn3 fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) return -1;
        fn = reinterpret_cast<EncodeTiledFn3>(sym);
    }
    cuuint64_t dims[3] = {(cuuint64_t)o.dim0, (cuuint64_t)o.dim1, 1};
    cuuint64_t strides[2] = {(cuuint64_t)o.stride1 * 2, (cuuint64_t)o.stride1 * o.dim1 * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(o.ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

template <int EPI>
static int g2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const Gemm2Params& p, cudaStream_t st) {
    auto kern = gemm2cta_kernel<EPI>;
    if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (size_t)g2_smem(EPI)) != A2V_OK) return A2V_ERR_CUDA;
    int pairs = a2v_num_sms() / 2;
    if (pairs > p.num_tiles) pairs = p.num_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(g2_threads(EPI));
    cfg.dynamicSmemBytes = g2_smem(EPI);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) {
        a2v_set_error("gemm(2cta): launch failed: %s", cudaGetErrorString(e));
        return A2V_ERR_CUDA;
    }
    return a2v_check_launch("gemm2cta_kernel");
}

static int g2_try_tn(const a2v_gemm_desc* d, cudaStream_t st) {
    if (d->taps != 1 || d->batch != 1 || d->groups != 1 || d->a_tap_cols != 0 || d->a_tap_wrap != 0 || d->c_dtype != A2V_F32) return -1;
    if (!d->out_atomic || d->M % 256 != 0 || d->N % 256 != 0 || d->ldc % 4 != 0 || d->red_rows < 64 * 16) return -1;
    if (d->a_row_off != 0 || d->b_row_off != 0 || d->c_row_off != 0) return -1;
    if (d->a.dim2 != 1 || d->b.dim2 != 1 || d->a.stride1 % 8 != 0 || d->b.stride1 % 8 != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(d->a.ptr) | reinterpret_cast<uintptr_t>(d->b.ptr) | reinterpret_cast<uintptr_t>(d->c)) & 15)
        return -1;
    Gemm2TnParams p;
    memset(&p, 0, sizeof(p));
    p.M = d->M; p.N = d->N;
    p.kblocks = ceil_div(d->red_rows, G2_BK);
    p.n_tiles = d->N / 256;
    p.m_tiles2 = d->M / 256;
    const int tiles = p.n_tiles * p.m_tiles2;
    const int pairs_total = a2v_num_sms() / 2;
    // split count (<= 16, at least 8 k-blocks each) that fills whole waves of CTA pairs best
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 16; ++s) {
        if (p.kblocks / s < 8 && s > 1) break;
        const int inst = tiles * s;
        const double eff = (double)inst / ((double)ceil_div(inst, pairs_total) * pairs_total);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    p.per_split = ceil_div(p.kblocks, best);
    p.splits = ceil_div(p.kblocks, p.per_split);
    p.num_tiles = tiles * p.splits;
    p.c = reinterpret_cast<float*>(d->c);
    p.ldc = d->ldc;
    p.alpha = d->alpha;
    CUtensorMap ta, tb;
    if (g2_make_map(&ta, d->a, 64) != 0 || g2_make_map(&tb, d->b, 64) != 0) return -1;
    auto kern = gemm2cta_tn_kernel;
    if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (size_t)G2_SMEM) != A2V_OK) return A2V_ERR_CUDA;
    int pairs = pairs_total < p.num_tiles ? pairs_total : p.num_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = G2_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) {
        a2v_set_error("gemm(2cta, tn): launch failed: %s", cudaGetErrorString(e));
        return A2V_ERR_CUDA;
    }
    return a2v_check_launch("gemm2cta_tn_kernel");
}

// Returns -1 when the descriptor is not a plain 2-D bf16 Linear the pair kernel handles (the caller then takes the
// one-CTA kernel), otherwise the launch status.
int gemm2cta_try(const a2v_gemm_desc* d, cudaStream_t st) {
    if (d->mode == 1) return g2_try_tn(d, st);
    if (d->mode != 0 || d->taps != 1 || d->batch != 1 || d->groups != 1 || d->c_dtype != A2V_BF16) return -1;
    if (d->out_atomic || d->out_accumulate) return -1;
    if (d->dgelu_u != nullptr && (d->bias || d->preact || d->residual || d->act != 0)) return -1;
    if (d->colsum != nullptr && d->dgelu_u == nullptr) return -1;
    if (d->block_n != 256 || d->N % G2_BN != 0 || d->k_per_tap % G2_BK != 0 || d->M < 4 * G2_BM) return -1;
    if (d->a_row_off != 0 || d->b_row_off != 0 || d->c_row_off != 0 || d->ldc % 8 != 0) return -1;
    if (d->a.dim2 != 1 || d->b.dim2 != 1 || d->a.stride1 % 8 != 0 || d->b.stride1 % 8 != 0) return -1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(d->a.ptr) | reinterpret_cast<uintptr_t>(d->b.ptr) |
                         reinterpret_cast<uintptr_t>(d->c) | reinterpret_cast<uintptr_t>(d->bias) |
                         reinterpret_cast<uintptr_t>(d->preact) | reinterpret_cast<uintptr_t>(d->residual) |
                         reinterpret_cast<uintptr_t>(d->dgelu_u);
    if (al & 15) return -1;
    const int m = (d->bias ? G2_EPI_BIAS : 0) | (d->preact ? G2_EPI_PREACT : 0) | (d->act == 1 ? G2_EPI_GELU : 0) |
                  (d->residual ? G2_EPI_RES : 0) | (d->dgelu_u ? G2_EPI_DGELU : 0);
    if (!(m == 0 || m == 1 || m == 5 || m == 7 || m == 8 || m == 16)) return -1;
    if (d->act != 0 && d->act != 1) return -1;
    Gemm2Params p;
    memset(&p, 0, sizeof(p));
    p.M = d->M; p.N = d->N; p.kblocks = d->k_per_tap / G2_BK;
    p.n_tiles = d->N / G2_BN;
    p.num_tiles = p.n_tiles * ceil_div(d->M, 2 * G2_BM);
    p.c = d->c; p.ldc = d->ldc; p.alpha = d->alpha; p.bias = d->bias; p.preact = d->preact; p.residual = d->dgelu_u ? d->dgelu_u : d->residual;
    p.colsum = d->colsum;
    CUtensorMap ta, tb;
    if (g2_make_map(&ta, d->a) != 0 || g2_make_map(&tb, d->b) != 0) return -1;
    switch (m) {
        case 0: return g2_launch<0>(ta, tb, p, st);
        case 1: return g2_launch<1>(ta, tb, p, st);
        case 5: return g2_launch<5>(ta, tb, p, st);
        case 7: return g2_launch<7>(ta, tb, p, st);
        case 8: return g2_launch<8>(ta, tb, p, st);
        default: return g2_launch<16>(ta, tb, p, st);
    }
}

}  // namespace a2v
