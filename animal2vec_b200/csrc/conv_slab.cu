// Grouped stride-1 "tap" convolution with 64-channel groups as a tcgen05 implicit GEMM that loads every
// activation row ONCE per tile (a row "slab" with the tap halo) instead of once per tap.
//
//   y[b, t, g*YG + n] = sum_{j < taps} sum_{c < 64} x[b, t + j - pad, g*64 + c] * w[g*WG + n, j*64 + c]  (+ bias)
//
// Replaces the cuDNN grouped Conv1d dispatch of the reference's relative positional encoder
// (/root/reference/nn/modalities/audio.py:93-113: 5 x Conv1d(D, D, k=19, groups=16)) and of Decoder1d
// (nn/modalities/modules.py:141-157: Conv1d(k=7, groups=16)), forward and data gradient (the data
// gradient is the same operator on dy with flipped/transposed weights, see params.pack_conv_dgrad).
//
// Why a slab: with one 128x64 A tile per tap (the generic tap loop of gemm_sm100.cu) a 19-tap layer pulls
// 19 x 16 KB of overlapping rows through L2 per 128x64 output tile (~190 B/clk/SM at the MMA rate): the
// kernel is L2-bandwidth bound. Here one TMA box brings rows [m0 - pad, m0 - pad + 256 + taps - 1) of
// the group's 64 channels into a 128B-swizzled slab; the A operand of tap j is the SAME shared memory
// shifted down by j rows (descriptor start address + j*128 B; the 128B swizzle is a function of the
// absolute shared-memory address, so the shifted view stays consistent with what TMA wrote). Two 128-row output tiles
// share each weight tile. Warp roles: warp 0 slab producer (4-stage ring: the activations stream from HBM once, the
// producer runs tiles ahead), warp 10 weight-tile producer (6-stage ring, L2 resident), warp 1 tcgen05.mma issuer (one
// thread), warps 2..9 epilogue out of double-buffered TMEM accumulators (one 128-row tile per warpgroup: with four
// epilogue warps the 7-tap layers were bound by the epilogue's serial chunk loop, not by the tensor pipe or HBM).
#include <string.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int CS_BLOCK_M = 128;
constexpr int CS_MT = 2;                        // 128-row tiles per slab
constexpr int CS_SUPER_M = CS_BLOCK_M * CS_MT;  // output rows per tile
constexpr int CS_NB = 6;                        // weight-tile ring stages (8 KB each)
constexpr int CS_NS = 3;                        // slab ring stages (a slab comes from HBM: the activations are streamed once)
constexpr int CS_B_BYTES = 64 * 64 * 2;
constexpr int CS_THREADS = 192;
constexpr int CS_FWD_THREADS = 352;             // forward: 8 epilogue warps (2..9, one 128-row tile each) + warp 10, the weight-tile producer
constexpr int CS_EPI_PITCH = 36;
constexpr int CS_EPI_BYTES = 4 * 32 * CS_EPI_PITCH * 4;
constexpr int CS_FWD_EPI_BYTES = 2 * CS_EPI_BYTES;  // forward: 8 epilogue warps

struct ConvSlabParams {
    int batch, T, groups, taps, pad, ng;
    int x_group_cols;  // channel distance between the groups of x (64; 48 for the compact decoder layout: the 64-wide box
                       // then overlaps the next group, whose channels only meet skipped K steps)
    int w_group_rows, y_group_cols;
    long long ldy;
    void* y;
    int y_f32;
    const float* bias;
    int slab_rows;   // multiple of 16, >= CS_SUPER_M + taps - 1
    int m_tiles;     // ceil(T / CS_SUPER_M)
    int num_tiles;
    int fast_store;  // bf16 output, 8-column groups 16-byte aligned: bf16 staging + 16-byte stores
};

__device__ __forceinline__ uint64_t umma_smem_desc_bo(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)(base_off & 7u) << 49;  // matrix base offset: start row inside the 8-row swizzle atom
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

template <typename TC>
__device__ __forceinline__ void cs_store_chunk(const ConvSlabParams& p, const float (&v)[32], long long row_off0,
                                               int rows_ok, int gcol, int cols_ok, int lane) {
    const int rsub = lane >> 3;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < cols_ok) bias[j] = p.bias[gcol + j];
    }
    TC* y = reinterpret_cast<TC*>(p.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 4 * i;
        if (r >= rows_ok || cols_ok <= 0) continue;
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = v[4 * i + j] + bias[j];
        TC* c = y + row_off0 + (long long)r * p.ldy + gcol;
        if (cols_ok >= 4) {
            store4(c, x);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < cols_ok) c[j] = from_f32<TC>(x[j]);
        }
    }
}

template <int KSTEPS>
__global__ void __launch_bounds__(CS_FWD_THREADS, 1)
conv_slab_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const ConvSlabParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slab_bytes = p.slab_rows * 128;
    uint8_t* slab0 = smem;
    uint8_t* bring = smem + CS_NS * slab_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bring + CS_NB * CS_B_BYTES);
    uint64_t* slab_full = bars;                 // [CS_NS]
    uint64_t* slab_empty = bars + CS_NS;        // [CS_NS]
    uint64_t* b_full = bars + 2 * CS_NS;        // [CS_NB]
    uint64_t* b_empty = b_full + CS_NB;    // [CS_NB]
    uint64_t* tfull = b_empty + CS_NB;     // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* epi_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < CS_NS; ++i) {
            mbar_init(&slab_full[i], 1);
            mbar_init(&slab_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 8);
        }
        for (int i = 0; i < CS_NB; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (group, m super tile, batch), group fastest: the CTAs running at the same time read (and write) ALL the
    // 128-byte group slices of the same activation rows, so every 2 KB row is consumed while its DRAM page is open and
    // its lines merge in L2 (group-major order streamed one 128-byte slice of every row per pass: ~50 % of the HBM
    // rate on the 7-tap layers). The weights of all groups (taps x 8 KB each) stay L2 resident either way.
    auto decode = [&](int tile, int& ms, int& b, int& g) {
        g = tile % p.groups;
        tile /= p.groups;
        ms = tile % p.m_tiles;
        b = tile / p.m_tiles;
    };

    if (warp == 0) {
        // slab producer: runs up to CS_NS tiles ahead of the tensor pipe, independent of the weight ring (one thread
        // for both kept the next slab behind the last weight tiles of the current one: ~6 taps of prefetch distance
        // against an HBM round trip)
        if (elect_one()) {
            int ss = 0;
            uint32_t sphase = 0;
            const int half_rows = p.slab_rows >> 1;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int ms, b, g;
                decode(tile, ms, b, g);
                mbar_wait(&slab_empty[ss], sphase ^ 1);
                uint8_t* slab = slab0 + ss * slab_bytes;
                mbar_expect_tx(&slab_full[ss], (uint32_t)slab_bytes);
                const int r0 = ms * CS_SUPER_M - p.pad;
                tma_load_3d(slab, &tmX, &slab_full[ss], g * p.x_group_cols, r0, b);
                tma_load_3d(slab + half_rows * 128, &tmX, &slab_full[ss], g * p.x_group_cols, r0 + half_rows, b);
                if (++ss == CS_NS) { ss = 0; sphase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // weight-tile producer
        if (elect_one()) {
            int bs = 0;
            uint32_t bphase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int ms, b, g;
                decode(tile, ms, b, g);
                for (int j = 0; j < p.taps; ++j) {
                    mbar_wait(&b_empty[bs], bphase ^ 1);
                    mbar_expect_tx(&b_full[bs], CS_B_BYTES);
                    tma_load_3d(bring + bs * CS_B_BYTES, &tmW, &b_full[bs], j * 64, g * p.w_group_rows, 0);
                    if (++bs == CS_NB) { bs = 0; bphase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(CS_BLOCK_M, 64, false, false);
            int ss = 0, bs = 0, as = 0;
            uint32_t sphase = 0, bphase = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                mbar_wait(&slab_full[ss], sphase);
                tc_fence_after();
                const uint32_t slab = smem_u32(slab0 + ss * slab_bytes);
                for (int j = 0; j < p.taps; ++j) {
                    mbar_wait(&b_full[bs], bphase);
                    tc_fence_after();
                    const uint32_t wb = smem_u32(bring + bs * CS_B_BYTES);
                    // measured on B200: the swizzle phase comes from the absolute smem address bits [7:9]; a row-shifted
                    // start address needs NO base-offset correction
                    const uint32_t bo = 0u;
#pragma unroll
                    for (int mt = 0; mt < CS_MT; ++mt) {
                        const uint32_t a0 = slab + (uint32_t)(mt * CS_BLOCK_M + j) * 128u;
                        const uint32_t td = tmem_base + as * (CS_MT * 64) + mt * 64;
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_bf16(td, umma_smem_desc_bo(a0 + k * 32, 1024, bo), umma_smem_desc(wb + k * 32, 0, 1024),
                                      idesc, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&b_empty[bs]);
                    if (++bs == CS_NB) { bs = 0; bphase ^= 1; }
                }
                umma_commit(&slab_empty[ss]);
                umma_commit(&tfull[as]);
                if (++ss == CS_NS) { ss = 0; sphase ^= 1; }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        int as = 0;
        uint32_t aphase = 0;
        const uint32_t st = smem_u32(epi_stage + (warp - 2) * (32 * CS_EPI_PITCH));
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int ms, b, g;
            decode(tile, ms, b, g);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            {
                const int mt = (warp - 2) >> 2;  // warps 2..5: rows 0..127 of the tile, warps 6..9: rows 128..255
                const int row0 = ms * CS_SUPER_M + mt * CS_BLOCK_M + q * 32;
                int rows_ok = p.T - row0;
                rows_ok = rows_ok > 32 ? 32 : rows_ok;
                const long long row_off0 = ((long long)b * p.T + row0) * p.ldy;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t raw[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (CS_MT * 64) + mt * 64 + c * 32, raw);
                    tmem_ld_wait();
                    if (c * 32 >= p.ng) break;
                    if (p.fast_store) {
                        // bf16 staging (half the shared-memory bytes of the fp32 staging below; this kernel is bound by
                        // shared-memory operand bandwidth): bias added in the TMEM-native layout (thread = row), packed,
                        // transposed through an 80-byte-pitch row, 16-byte stores (8 rows x 64 contiguous bytes each)
                        const uint32_t st16 = smem_u32(reinterpret_cast<uint8_t*>(epi_stage) + (warp - 2) * (32 * 80));
                        const float* bp = p.bias != nullptr ? p.bias + g * p.y_group_cols + c * 32 : nullptr;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(raw[8 * j + i]);
                            if (bp != nullptr && c * 32 + 8 * j < p.ng) {
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + 8 * j));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bp + 8 * j) + 1);
                                x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                                x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st16 + lane * 80 + 16 * j),
                                         "r"(pack_bf16x2(x[0], x[1])), "r"(pack_bf16x2(x[2], x[3])),
                                         "r"(pack_bf16x2(x[4], x[5])), "r"(pack_bf16x2(x[6], x[7]))
                                         : "memory");
                        }
                        __syncwarp();
                        const int rrow = lane & 7, rchunk = lane >> 3;
                        bf16* yb = reinterpret_cast<bf16*>(p.y) + row_off0 + g * p.y_group_cols + c * 32 + 8 * rchunk;
                        const bool col_ok = c * 32 + 8 * rchunk + 8 <= p.ng;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = rrow + 8 * i;
                            uint4 v4;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w)
                                         : "r"(st16 + r * 80 + 16 * rchunk)
                                         : "memory");
                            if (r < rows_ok && col_ok) *reinterpret_cast<uint4*>(yb + (long long)r * p.ldy) = v4;
                        }
                        __syncwarp();
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + (lane * CS_EPI_PITCH + 4 * j) * 4),
                                     "r"(raw[4 * j]), "r"(raw[4 * j + 1]), "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3])
                                     : "memory");
                    __syncwarp();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                                     : "r"(st + (((lane >> 3) + 4 * i) * CS_EPI_PITCH + (lane & 7) * 4) * 4)
                                     : "memory");
                    __syncwarp();
                    const int lcol = c * 32 + (lane & 7) * 4;
                    const int cols_ok = p.ng - lcol;
                    if (rows_ok > 0) {
                        if (p.y_f32)
                            cs_store_chunk<float>(p, v, row_off0, rows_ok, g * p.y_group_cols + lcol, cols_ok, lane);
                        else
                            cs_store_chunk<bf16>(p, v, row_off0, rows_ok, g * p.y_group_cols + lcol, cols_ok, lane);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of the same operator, x as the MMA M side with the taps folded into M:
//   dW_T[(g*taps + j)*64 + c, n] += sum_{b,t} x[b, t + j - pad, g*64 + c] * dy[b, t, g*DG + n]
// One CTA owns (group, block of 16 taps, K split). Per 64-row k-block it loads ONE x slab (64 + 15 halo
// rows) and one dy tile; M tile i pairs tap i with tap i+8 (the second 64-wide M chunk is the same slab
// 8 rows = 1024 B further down, so the descriptor's leading-dimension byte offset is 1024 and the swizzle
// phase is unchanged). Up to 8 accumulators (512 TMEM columns) stay resident over the whole K range; the
// epilogue adds them to the fp32 gradient with atomics (split-K).
// ---------------------------------------------------------------------------------------------
constexpr int CW_STAGES = 8;
constexpr int CW_SLAB_ROWS = 80;                       // 64 + 15 halo, rounded to 8
constexpr int CW_SLAB_BYTES = CW_SLAB_ROWS * 128;      // 10 KB
constexpr int CW_DY_BYTES = 64 * 128;                  // 8 KB
constexpr int CW_STAGE_BYTES = CW_SLAB_BYTES + CW_DY_BYTES;

struct ConvWgradParams {
    int batch, T, groups, taps, pad, ng;
    int x_group_cols;  // as in ConvSlabParams; the M rows of the overlap (channels >= x_group_cols) are scratch rows of `out`
    int dy_group_cols;
    float* out;          // (groups*taps*64, ldo) fp32
    long long ldo;
    int kb_per_batch;    // ceil(T / 64)
    int n_tb;            // tap blocks (<= 2)
    int nt[2];           // M tiles per tap block
    int splits[2];       // K splits per tap block
    int num_tiles;
    // 5..8 taps: a SECOND copy of the x slab, loaded 4 rows further down, is the second 64-wide M chunk of every tile, so
    // M tile i pairs tap i with tap i + 4 (pairing with tap i + 8 -- the same slab 8 rows down -- would leave half of
    // every 128-row MMA on taps that do not exist: the 7-tap decoder layers ran at half rate)
    int pair_off;        // 8, or 4 with the second slab copy
    int stages;          // pipeline stages (8, or 6 with the second slab copy)
    int stage_bytes;
};

__global__ void __launch_bounds__(CS_THREADS, 1)
conv_slab_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                       const ConvWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + CW_STAGES;
    uint64_t* tfull = empty_bar + CW_STAGES;  // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float* epi_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDy);
        for (int i = 0; i < CW_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // one tile per CTA (grid == num_tiles): tile -> (tap block, group, split)
    int tile = blockIdx.x;
    int tb = 0;
    if (tile >= p.groups * p.splits[0]) {
        tile -= p.groups * p.splits[0];
        tb = 1;
    }
    const int g = tile % p.groups;
    const int split = tile / p.groups;
    const int nt = p.nt[tb];
    const int total_kb = p.batch * p.kb_per_batch;
    const int per = (total_kb + p.splits[tb] - 1) / p.splits[tb];
    const int kb0 = split * per;
    const int kb1 = min(total_kb, kb0 + per);

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int b = kb / p.kb_per_batch;
                const int r0 = (kb - b * p.kb_per_batch) * 64;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sx = smem + stage * p.stage_bytes;
                mbar_expect_tx(&full_bar[stage], (uint32_t)p.stage_bytes);
                tma_load_3d(sx, &tmX, &full_bar[stage], g * p.x_group_cols, r0 + tb * 16 - p.pad, b);
                if (p.pair_off == 4)
                    tma_load_3d(sx + CW_SLAB_BYTES, &tmX, &full_bar[stage], g * p.x_group_cols, r0 + 4 - p.pad, b);
                tma_load_3d(sx + p.stage_bytes - CW_DY_BYTES, &tmDy, &full_bar[stage], g * p.dy_group_cols, r0, b);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, 64, true, true);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sx = smem_u32(smem + stage * p.stage_bytes);
                const uint32_t sd = sx + (uint32_t)(p.stage_bytes - CW_DY_BYTES);
                // A: M-major, 16 K rows per step (2048 B); M chunk 1 = 8 rows (1024 B) below chunk 0, or the same rows of
                // the second slab copy (which starts 4 x rows later)
                const uint32_t lbo = p.pair_off == 4 ? (uint32_t)CW_SLAB_BYTES : 1024u;
                for (int i = 0; i < nt; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = umma_smem_desc(sx + (uint32_t)i * 128u + k * 2048, lbo, 1024);
                        const uint64_t db = umma_smem_desc(sd + k * 2048, 0, 1024);
                        umma_bf16(tmem_base + i * 64, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull);
        }
        __syncwarp();
    } else if (kb1 > kb0) {
        const int q = warp & 3;
        const uint32_t st = smem_u32(epi_stage + (warp - 2) * (32 * CS_EPI_PITCH));
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < nt; ++i) {
            const int tap = tb * 16 + i + (q >= 2 ? p.pair_off : 0);
            const int ch0 = (q & 1) * 32;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + i * 64 + c * 32, raw);
                tmem_ld_wait();
                if (tap >= p.taps || c * 32 >= p.ng) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + (lane * CS_EPI_PITCH + 4 * j) * 4),
                                 "r"(raw[4 * j]), "r"(raw[4 * j + 1]), "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3])
                                 : "memory");
                __syncwarp();
                float v[32];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[4 * r]), "=f"(v[4 * r + 1]), "=f"(v[4 * r + 2]), "=f"(v[4 * r + 3])
                                 : "r"(st + (((lane >> 3) + 4 * r) * CS_EPI_PITCH + (lane & 7) * 4) * 4)
                                 : "memory");
                __syncwarp();
                const int lcol = c * 32 + (lane & 7) * 4;
                float* orow = p.out + ((long long)(g * p.taps + tap) * 64 + ch0) * p.ldo + lcol;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float* o = orow + (long long)((lane >> 3) + 4 * r) * p.ldo;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (lcol + j < p.ng) atomicAdd(o + j, v[4 * r + j]);
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn3 cs_encode_fn() {
    static EncodeTiledFn3 fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn3>(sym);
    }
    return fn;
}

static int cs_make_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long batch, long long ld,
                       int box_rows, const char* name) {
    EncodeTiledFn3 fn = cs_encode_fn();
    if (fn == nullptr) {
        a2v_set_error("conv_slab: cuTensorMapEncodeTiled not available");
        return A2V_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * rows * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        a2v_set_error("conv_slab: cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
        return A2V_ERR_CUDA;
    }
    return A2V_OK;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_conv_slab_supported(const a2v_conv_desc* d) {
    // x_group_cols < 64 (compact group-padded layouts): the 64-channel box overlaps the next group, so the caller must
    // declare x_real_cols <= x_group_cols -- the K steps over the overlap are then never issued (forward), or land in
    // scratch rows of the output (weight gradient)
    return d != nullptr && d->taps >= 1 && d->taps <= 32 && d->ng >= 1 && d->ng <= 64 && d->ldx % 8 == 0 &&
           (d->x_group_cols == 64 || (d->x_group_cols >= 16 && d->x_group_cols < 64 && d->x_group_cols % 16 == 0 &&
                                      d->x_real_cols > 0 && d->x_real_cols <= d->x_group_cols)) &&
           d->T >= 1 && d->pad >= 0 && d->pad < d->taps;
}

extern "C" int a2v_conv_slab_fwd(const a2v_conv_desc* d, a2v_stream_t stream) {
    A2V_REQUIRE(d != nullptr && d->x && d->w && d->y, "conv_slab: NULL pointer");
    A2V_REQUIRE(a2v_conv_slab_supported(d), "conv_slab: needs 64-channel groups, ng <= 64, taps <= 32");
    A2V_REQUIRE(d->y_dtype == A2V_F32 || d->y_dtype == A2V_BF16, "conv_slab: bad y dtype");
    A2V_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w & 15) == 0 && ((uintptr_t)d->y & 15) == 0,
                "conv_slab: pointers must be 16-byte aligned");
    const int vec = d->y_dtype == A2V_F32 ? 4 : 8;
    A2V_REQUIRE(d->ldy % vec == 0 && d->y_group_cols % vec == 0, "conv_slab: ldy / group stride alignment");
    A2V_REQUIRE(d->ldw % 8 == 0 && d->ldw >= (int64_t)d->taps * 64, "conv_slab: bad weight row stride");
    ConvSlabParams p;
    memset(&p, 0, sizeof(p));
    p.batch = d->batch; p.T = d->T; p.groups = d->groups; p.taps = d->taps; p.pad = d->pad; p.ng = d->ng;
    p.x_group_cols = d->x_group_cols; p.w_group_rows = d->w_group_rows; p.y_group_cols = d->y_group_cols;
    p.ldy = d->ldy; p.y = d->y; p.y_f32 = d->y_dtype == A2V_F32; p.bias = d->bias;
    p.slab_rows = (CS_SUPER_M + d->taps - 1 + 15) & ~15;
    p.m_tiles = ceil_div(d->T, CS_SUPER_M);
    p.num_tiles = p.m_tiles * d->batch * d->groups;
    A2V_REQUIRE(d->x_real_cols >= 0 && d->x_real_cols <= 64, "conv_slab: x_real_cols out of range");
    // K steps of 16 channels per tap: 3 when the last 16 channels of every input group are zero padding, else 4
    const int ksteps = (d->x_real_cols > 0 && d->x_real_cols <= 48) ? 3 : 4;
    A2V_REQUIRE(d->x_group_cols == 64 || ksteps * 16 <= d->x_group_cols, "conv_slab: compact groups need K steps inside the group (x_group_cols %d, %d steps)", d->x_group_cols, ksteps);
    p.fast_store = !p.y_f32 && d->ng % 8 == 0 && d->ldy % 8 == 0 && d->y_group_cols % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(d->y) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0;
    CUtensorMap tx, tw;
    int rc;
    if ((rc = cs_make_map(&tx, d->x, d->ldx, d->T, d->batch, d->ldx, p.slab_rows / 2, "x")) != A2V_OK) return rc;
    if ((rc = cs_make_map(&tw, d->w, d->ldw, (long long)d->groups * d->w_group_rows, 1, d->ldw, 64, "w")) != A2V_OK)
        return rc;
    const int smem = CS_NS * p.slab_rows * 128 + CS_NB * CS_B_BYTES + 256 + CS_FWD_EPI_BYTES + 1024;
    const int grid = p.num_tiles < a2v_num_sms() ? p.num_tiles : a2v_num_sms();
    if (ksteps == 3) {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(conv_slab_fwd_kernel<3>), (size_t)smem) != A2V_OK) return A2V_ERR_CUDA;
        conv_slab_fwd_kernel<3><<<grid, CS_FWD_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, p);
    } else {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(conv_slab_fwd_kernel<4>), (size_t)smem) != A2V_OK) return A2V_ERR_CUDA;
        conv_slab_fwd_kernel<4><<<grid, CS_FWD_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, p);
    }
    return a2v_check_launch("conv_slab_fwd");
}

extern "C" int a2v_conv_slab_wgrad(const a2v_conv_desc* d, float* out, int64_t ldo, a2v_stream_t stream) {
    // d->x: activations (batch, T, ldx); d->w: dy (batch, T, ldw) bf16 with w_group_rows = columns of dy per
    // group; out: (groups*taps*64, ldo) fp32, atomically accumulated
    A2V_REQUIRE(d != nullptr && d->x && d->w && out, "conv_slab_wgrad: NULL pointer");
    A2V_REQUIRE(a2v_conv_slab_supported(d), "conv_slab_wgrad: needs 64-channel groups, ng <= 64, taps <= 32");
    A2V_REQUIRE(d->ldw % 8 == 0 && ldo >= d->ng, "conv_slab_wgrad: bad strides");
    ConvWgradParams p;
    memset(&p, 0, sizeof(p));
    p.batch = d->batch; p.T = d->T; p.groups = d->groups; p.taps = d->taps; p.pad = d->pad; p.ng = d->ng;
    p.x_group_cols = d->x_group_cols;
    p.dy_group_cols = d->w_group_rows;
    p.out = out; p.ldo = ldo;
    p.kb_per_batch = ceil_div(d->T, 64);
    p.n_tb = ceil_div(d->taps, 16);
    const bool narrow = d->taps >= 5 && d->taps <= 8;
    p.pair_off = narrow ? 4 : 8;
    p.stages = narrow ? 6 : CW_STAGES;
    p.stage_bytes = (narrow ? 2 : 1) * CW_SLAB_BYTES + CW_DY_BYTES;
    for (int tb = 0; tb < 2; ++tb) {
        const int rem = d->taps - 16 * tb;
        p.nt[tb] = tb < p.n_tb ? (rem < p.pair_off ? rem : p.pair_off) : 0;
        p.splits[tb] = 0;
    }
    // K splits per tap block: fill the SMs once, minimise the longest CTA (work ~ M tiles / splits)
    const int sms = a2v_num_sms();
    const int total_kb = p.batch * p.kb_per_batch;
    int max_s = sms / d->groups;
    if (max_s < p.n_tb) max_s = p.n_tb;
    if (p.n_tb == 1) {
        p.splits[0] = max_s < total_kb ? max_s : total_kb;
    } else {
        double best = 1e30;
        for (int s0 = 1; s0 < max_s; ++s0) {
            const int s1 = max_s - s0;
            const double w0 = (double)p.nt[0] / s0, w1 = (double)p.nt[1] / s1;
            const double w = w0 > w1 ? w0 : w1;
            if (w < best) { best = w; p.splits[0] = s0; p.splits[1] = s1; }
        }
        if (p.splits[0] > total_kb) p.splits[0] = total_kb;
        if (p.splits[1] > total_kb) p.splits[1] = total_kb;
    }
    if (p.splits[0] < 1) p.splits[0] = 1;
    if (p.n_tb > 1 && p.splits[1] < 1) p.splits[1] = 1;
    p.num_tiles = d->groups * (p.splits[0] + p.splits[1]);
    CUtensorMap tx, tdy;
    int rc;
    if ((rc = cs_make_map(&tx, d->x, d->ldx, d->T, d->batch, d->ldx, CW_SLAB_ROWS, "x")) != A2V_OK) return rc;
    if ((rc = cs_make_map(&tdy, d->w, d->ldw, d->T, d->batch, d->ldw, 64, "dy")) != A2V_OK) return rc;
    const int smem = p.stages * p.stage_bytes + 256 + CS_EPI_BYTES + 1024;
    if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(conv_slab_wgrad_kernel), (size_t)smem) != A2V_OK) return A2V_ERR_CUDA;
    conv_slab_wgrad_kernel<<<p.num_tiles, CS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tdy, p);
    return a2v_check_launch("conv_slab_wgrad");
}
