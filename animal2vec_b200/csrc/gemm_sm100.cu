// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0      : TMA producer (cp.async.bulk.tensor into a 128B-swizzled smem ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> bias/GELU/residual -> global)
//
// The accumulator lives in TMEM (two stages of BLOCK_N fp32 columns) so the epilogue of
// tile i overlaps the MMAs of tile i+1. Operands are bf16, the accumulator fp32.
// Two operand modes (see include/a2v_capi.h):
//   NT : A, B K-major.  Tap loop over shifted A rows = stride-1 (grouped) conv1d.
//   TN : A, B MN-major. Reduction over (batch, rows) = weight gradients, optional split-K.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;       // 2 control warps + 4 epilogue warps
constexpr int GEMM_THREADS_WIDE = 320;  // GELU epilogues: 8 epilogue warps (two per TMEM lane quarter, half the columns each)

struct GemmParams {
    int mode;
    int M, N;
    int kb_per_tap, taps, batch, groups;
    int m_tiles, n_tiles;
    int a_group_stride, a_row_off, a_tap_rows, a_tap_cols;
    int a_tap_wrap, a_grow_add, a_grow_div, a_tap_col_stride;  // strided / gathered conv addressing (see a2v_gemm_desc)
    int b_group_stride, b_row_off, b_tap_rows;
    int red_rows, kb_per_batch, k_splits;
    void* c;
    int c_f32, out_atomic, out_accumulate;
    long long ldc, c_batch_stride, c_row_off;
    int c_group_stride, c_tap_stride;
    float alpha;
    const float* bias;
    int act;
    void* preact;
    const void* residual;
    const void* dgelu_u;
    int num_tiles;
};

struct TileCoord {
    int n_tile, m_tile, b, tap, g, split;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
    TileCoord t;
    t.n_tile = tile % p.n_tiles;
    tile /= p.n_tiles;
    t.m_tile = tile % p.m_tiles;
    tile /= p.m_tiles;
    if (p.mode == 0) {
        t.b = tile % p.batch;
        t.g = tile / p.batch;
        t.tap = 0;
        t.split = 0;
    } else {
        t.tap = tile % p.taps;
        tile /= p.taps;
        t.g = tile % p.groups;
        t.split = tile / p.groups;
        t.b = 0;
    }
    return t;
}

// k-block range [kb0, kb1) a tile iterates over
__device__ __forceinline__ void tile_k_range(const GemmParams& p, const TileCoord& t, int& kb0, int& kb1) {
    if (p.mode == 0) {
        kb0 = 0;
        kb1 = p.taps * p.kb_per_tap;
    } else {
        const int total = p.batch * p.kb_per_batch;
        const int per = (total + p.k_splits - 1) / p.k_splits;
        kb0 = t.split * per;
        kb1 = min(total, kb0 + per);
        if (kb1 < kb0) kb1 = kb0;
    }
}

constexpr int EPI_WIDE_PITCH = 80;  // bytes per staged bf16 row (64 + 16): conflict-free 16-byte accesses both ways
template <int BLOCK_N, bool WIDE = false>
struct GemmSmem {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int EPI_BYTES = WIDE ? 8 * 32 * EPI_WIDE_PITCH   // 8 epilogue warps x 32 rows x 80 B (bf16 staging)
                                          : 4 * 32 * 36 * 4;          // 4 epilogue warps x 32 rows x EPI_PITCH floats
    // per-tile bias vector staged per epilogue warp BEFORE it waits for the accumulator (a global load inside the
    // chunk loop costs a full memory latency per 32-column chunk, which a K = 1024 tile cannot hide)
    static constexpr int BIAS_BYTES = WIDE ? 8 * (BLOCK_N / 2) * 4 : 4 * BLOCK_N * 4;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + BIAS_BYTES + 1024;  // +1024: alignment slack
    static constexpr int TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
};

template <typename T>
__device__ __forceinline__ void load32(const T* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t[4];
        load4(p + 4 * i, t);
        v[4 * i] = t[0]; v[4 * i + 1] = t[1]; v[4 * i + 2] = t[2]; v[4 * i + 3] = t[3];
    }
}
template <typename T>
__device__ __forceinline__ void store32(T* p, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
        store4(p + 4 * i, t);
    }
}

// Epilogue for one 32-column chunk of 32 accumulator rows, in the COALESCED layout produced by the
// shared-memory transpose: lane l holds columns 4*(l&7) .. +3 of rows (l>>3) + 4*i, i = 0..7 in
// v[4*i .. 4*i+3]. A warp-level access therefore touches 4 rows x 128 contiguous bytes (fp32) or
// 4 rows x 64 bytes (bf16): full sectors, 4 lines per instruction instead of 32.
constexpr int EPI_PITCH = 36;  // floats per staged row (32 + 4): conflict-free for both access patterns

// Compile-time epilogue selection: EPI >= 0 is a bit mask (the hot variants, straight-line code that
// fits the instruction cache); EPI < 0 reads the descriptor at run time (every other combination).
constexpr int EPI_BIAS = 1, EPI_PREACT = 2, EPI_GELU = 4, EPI_DGELU = 8, EPI_RES = 16;
template <int EPI> __device__ __forceinline__ bool epi_bias(const GemmParams& p) { return EPI < 0 ? p.bias != nullptr : (EPI & EPI_BIAS) != 0; }
template <int EPI> __device__ __forceinline__ bool epi_preact(const GemmParams& p) { return EPI < 0 ? p.preact != nullptr : (EPI & EPI_PREACT) != 0; }
template <int EPI> __device__ __forceinline__ bool epi_gelu(const GemmParams& p) { return EPI < 0 ? p.act == 1 : (EPI & EPI_GELU) != 0; }
template <int EPI> __device__ __forceinline__ bool epi_dgelu(const GemmParams& p) { return EPI < 0 ? p.dgelu_u != nullptr : (EPI & EPI_DGELU) != 0; }
template <int EPI> __device__ __forceinline__ bool epi_res(const GemmParams& p) { return EPI < 0 ? p.residual != nullptr : (EPI & EPI_RES) != 0; }

template <typename TC, int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&v)[32], long long row_off0, int rows_ok,
                                               int gcol, int cols_ok, int lane, const float* bias_s = nullptr) {
    // row_off0: element offset of the warp's first row; rows_ok: number of valid rows (<= 32) from it;
    // gcol: global column of this lane's first element; cols_ok: valid columns from gcol (may be <= 0)
    const int rsub = lane >> 3;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = cols_ok >= 4;
    if (epi_bias<EPI>(p) && cols_ok > 0) {
        if (bias_s != nullptr) {  // staged per tile in shared memory (pads are zeros)
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s);
            bias[0] = b4.x; bias[1] = b4.y; bias[2] = b4.z; bias[3] = b4.w;
        } else if (full) {
            load4(p.bias + gcol, bias);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cols_ok) bias[j] = p.bias[gcol + j];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 4 * i;
        if (r >= rows_ok || cols_ok <= 0) continue;
        const long long off = row_off0 + (long long)r * p.ldc + gcol;
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = v[4 * i + j] * p.alpha + bias[j];
        if (epi_preact<EPI>(p)) {
            TC* u = reinterpret_cast<TC*>(p.preact) + off;
            if (full) store4(u, x);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < cols_ok) u[j] = from_f32<TC>(x[j]);
            }
        }
        if (epi_gelu<EPI>(p)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = gelu_t<TC>(x[j]);
        }
        if (epi_dgelu<EPI>(p)) {
            const TC* u = reinterpret_cast<const TC*>(p.dgelu_u) + off;
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            if (full) load4(u, t);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < cols_ok) t[j] = to_f32(u[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] *= gelu_grad_t<TC>(t[j]);
        }
        if (epi_res<EPI>(p)) {
            const TC* rr = reinterpret_cast<const TC*>(p.residual) + off;
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            if (full) load4(rr, t);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < cols_ok) t[j] = to_f32(rr[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] += t[j];
        }
        TC* c = reinterpret_cast<TC*>(p.c) + off;
        if (full) store4(c, x);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cols_ok) c[j] = from_f32<TC>(x[j]);
        }
    }
}

__device__ __forceinline__ void epilogue_chunk_f32_accum(const GemmParams& p, float (&v)[32], long long row_off0,
                                                         int rows_ok, int gcol, int cols_ok, int lane) {
    const int rsub = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 4 * i;
        if (r >= rows_ok || cols_ok <= 0) continue;
        float* c = reinterpret_cast<float*>(p.c) + row_off0 + (long long)r * p.ldc + gcol;
        if (p.out_atomic) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cols_ok) atomicAdd(c + j, v[4 * i + j] * p.alpha);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cols_ok) c[j] += v[4 * i + j] * p.alpha;
        }
    }
}

// The 8-warp epilogue with bf16 staging serves the compile-time variants plain, bias, bias+GELU(+preact) and +residual
// (the residual rows are prefetched in the TMEM-native layout, thread = row, before the accumulator wait): its shared-memory staging moves half the bytes of the fp32 staging of the 4-warp epilogue, and
// shared-memory bandwidth is what the epilogue competes for with the MMA operand reads (ncu: the epilogue's stores wait
// on the short scoreboard of the staged LDS while tcgen05.mma streams 96 B/clk of operands).
template <int EPI> constexpr bool epi_is_wide() { return EPI >= 0 && (EPI & EPI_DGELU) == 0; }

template <int BLOCK_N, int MODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS_WIDE, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
    constexpr bool WIDE = epi_is_wide<EPI>();
    using S = GemmSmem<BLOCK_N, WIDE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* tfull_bar = empty_bar + S::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* epi_stage = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + S::BAR_BYTES);
    float* bias_stage = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + S::BAR_BYTES + S::EPI_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < S::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], WIDE ? 8 : 4);
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        tmem_alloc<S::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                int kb0, kb1;
                tile_k_range(p, t, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    uint8_t* sb = sa + S::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                    if (MODE == 0) {
                        const int tap = kb / p.kb_per_tap;
                        const int kc = kb - tap * p.kb_per_tap;
                        int arow = t.m_tile * BLOCK_M + tap * p.a_tap_rows + p.a_row_off;
                        int acol = t.g * p.a_group_stride + kc * BLOCK_K;
                        if (p.a_tap_wrap > 0) {  // strided conv: (B, T, C) viewed as (B, T/s, s*C)
                            const int q = tap + p.a_row_off;
                            const int rq = q >= 0 ? q / p.a_tap_wrap : -((-q + p.a_tap_wrap - 1) / p.a_tap_wrap);
                            arow = t.m_tile * BLOCK_M + rq;
                            acol += (q - rq * p.a_tap_wrap) * p.a_tap_col_stride;
                        }
                        if (p.a_grow_div > 0) arow += (t.g + p.a_grow_add) / p.a_grow_div;
                        tma_load_3d(sa, &tmA, &full_bar[stage], acol, arow, t.b);
                        tma_load_3d(sb, &tmB, &full_bar[stage], kb * BLOCK_K,
                                    t.g * p.b_group_stride + t.n_tile * BLOCK_N, 0);
                    } else {
                        const int b = kb / p.kb_per_batch;
                        const int r0 = (kb - b * p.kb_per_batch) * BLOCK_K;
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i) {
                            // a_tap_cols > 0: the M index is (tap, channel); every 64-column box of A reads
                            // rows shifted by its own tap (weight gradient of a stride-1 conv, x as the M side)
                            const int m0 = t.m_tile * BLOCK_M + i * 64;
                            const int tap = p.a_tap_cols > 0 ? m0 / p.a_tap_cols : 0;
                            const int c0 = m0 - tap * p.a_tap_cols;
                            int arow = r0 + (p.a_tap_cols > 0 ? tap * p.a_tap_rows + p.a_row_off : 0);
                            int acol = t.g * p.a_group_stride + c0;
                            if (p.a_tap_wrap > 0) {  // strided conv weight gradient: x viewed as (B, T/s, s*C)
                                const int q = tap + p.a_row_off;
                                const int rq = q >= 0 ? q / p.a_tap_wrap : -((-q + p.a_tap_wrap - 1) / p.a_tap_wrap);
                                arow = r0 + rq;
                                acol += (q - rq * p.a_tap_wrap) * p.a_tap_col_stride;
                            }
                            tma_load_3d(sa + i * (64 * BLOCK_K * 2), &tmA, &full_bar[stage], acol, arow, b);
                        }
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i)
                            tma_load_3d(sb + i * (64 * BLOCK_K * 2), &tmB, &full_bar[stage],
                                        t.g * p.b_group_stride + t.n_tile * BLOCK_N + i * 64,
                                        r0 + t.tap * p.b_tap_rows + p.b_row_off, b);
                    }
                    if (++stage == S::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, MODE == 1, MODE == 1);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                int kb0, kb1;
                tile_k_range(p, t, kb0, kb1);
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint32_t b_base = a_base + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        uint64_t da, db;
                        if (MODE == 0) {
                            da = umma_smem_desc(a_base + k * (UMMA_K * 2), 0, 1024);
                            db = umma_smem_desc(b_base + k * (UMMA_K * 2), 0, 1024);
                        } else {
                            da = umma_smem_desc(a_base + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024);
                            db = umma_smem_desc(b_base + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024);
                        }
                        umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == S::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[as]);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (WIDE) {
        // ------------------------------------------------------------ epilogue warps, GELU variants
        // Eight warps: warp w reads TMEM lane quarter w % 4 and half of the tile's 32-column chunks. The
        // arithmetic runs in the TMEM-native layout (thread = row; the bias vector is a broadcast load), the
        // results are packed to bf16, transposed through an 80-byte-pitch staging row and leave as 16-byte
        // stores (8 rows x 64 contiguous bytes per instruction).
        const int q = warp & 3;
        const int ew = warp - 2;
        constexpr int CPW = (BLOCK_N / 32) / 2 > 0 ? (BLOCK_N / 32) / 2 : 1;  // chunks per warp
        const int c_begin = (ew >> 2) * CPW;
        const uint32_t st = smem_u32(reinterpret_cast<uint8_t*>(epi_stage) + ew * (32 * EPI_WIDE_PITCH));
        const int rrow = lane & 7, rchunk = lane >> 3;  // staged read: row rrow + 8 i, 16-byte chunk rchunk
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            const int row0_in_block = t.m_tile * BLOCK_M + q * 32;
            int rows_ok = p.M - row0_in_block;
            rows_ok = rows_ok > 32 ? 32 : rows_ok;
            const long long row_g = (long long)t.b * p.c_batch_stride + p.c_row_off + row0_in_block;
            const int col_base = t.g * p.c_group_stride;
            const long long row_off0 = row_g * p.ldc;
            float* bs = bias_stage + ew * (CPW * 32);  // this warp's CPW chunks of the tile's bias vector
            if (EPI & EPI_BIAS) {
                const int cbase = t.n_tile * BLOCK_N + c_begin * 32;
                const float* bg = p.bias + col_base + cbase;
                const int ncols = p.N - cbase;
                for (int j = lane * 4; j < CPW * 32; j += 128) {
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j + 4 <= ncols) b4 = __ldg(reinterpret_cast<const float4*>(bg + j));
                    *reinterpret_cast<float4*>(bs + j) = b4;
                }
                __syncwarp();
            }
            // residual variant: this thread's row of the residual tile (64 bytes per 32-column chunk), every load of the
            // tile issued NOW so that their latency hides behind the MMAs of the tile
            constexpr bool RES = (EPI & EPI_RES) != 0;
            uint4 rsd[RES ? CPW : 1][4];
            if (RES) {
                const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + row_off0 + (long long)lane * p.ldc + col_base;
#pragma unroll
                for (int cc = 0; cc < CPW; ++cc) {
                    const int col0 = t.n_tile * BLOCK_N + (c_begin + cc) * 32;
                    const bool ok = lane < rows_ok && (c_begin + cc) < BLOCK_N / 32 && col0 < p.N;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rsd[RES ? cc : 0][j] = ok ? __ldg(reinterpret_cast<const uint4*>(rp + col0) + j) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
#pragma unroll(RES ? CPW : 1)
            for (int cc = 0; cc < CPW; ++cc) {
                const int c = c_begin + cc;
                const int col0 = t.n_tile * BLOCK_N + c * 32;
                if (c >= BLOCK_N / 32 || col0 >= p.N) break;
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N + c * 32, raw);
                tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bv = (EPI & EPI_BIAS) ? *reinterpret_cast<const float4*>(bs + cc * 32 + 4 * j)
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
                        x[8 * j] += r0.x; x[8 * j + 1] += r0.y; x[8 * j + 2] += r1.x; x[8 * j + 3] += r1.y;
                        x[8 * j + 4] += r2.x; x[8 * j + 5] += r2.y; x[8 * j + 6] += r3.x; x[8 * j + 7] += r3.y;
                    }
                }
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 0 && !(EPI & EPI_PREACT)) continue;
                    bf16* dst = reinterpret_cast<bf16*>(pass == 0 ? p.preact : p.c);
                    if (pass == 1 && (EPI & EPI_GELU)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) x[i] = gelu_tanh_fast(x[i]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w0 = pack_bf16x2(x[8 * j], x[8 * j + 1]), w1 = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
                        const uint32_t w2 = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), w3 = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + lane * EPI_WIDE_PITCH + 16 * j),
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
                                     : "r"(st + r * EPI_WIDE_PITCH + 16 * rchunk)
                                     : "memory");
                        if (r < rows_ok)
                            *reinterpret_cast<uint4*>(dst + row_off0 + (long long)r * p.ldc + col_base + col0 + 8 * rchunk) = v;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------ epilogue warps
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            int kb0, kb1;
            tile_k_range(p, t, kb0, kb1);
            const int row0_in_block = t.m_tile * BLOCK_M + q * 32;  // first row of this warp's 32-row slab
            int rows_ok = p.M - row0_in_block;
            rows_ok = rows_ok > 32 ? 32 : rows_ok;
            long long row_g;
            int col_base;
            if (MODE == 0) {
                row_g = (long long)t.b * p.c_batch_stride + p.c_row_off + row0_in_block;
                col_base = t.g * p.c_group_stride;
            } else {
                row_g = (long long)t.g * p.c_group_stride + row0_in_block;
                col_base = t.tap * p.c_tap_stride;
            }
            const long long row_off0 = row_g * p.ldc;
            const bool has_work = (kb1 > kb0);
            const uint32_t st = smem_u32(epi_stage + (warp - 2) * (32 * EPI_PITCH));
            // residual epilogue: a load issued inside the chunk loop costs a full memory latency per chunk
            // (measured 4x the GEMM itself); issue every load of the tile NOW, before waiting for the MMAs
            constexpr bool PREFETCH = (EPI == EPI_RES);
            uint2 aux[PREFETCH ? BLOCK_N / 32 : 1][8];
            if (PREFETCH) {
                const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + row_off0 + col_base;
#pragma unroll
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    const int lcol = t.n_tile * BLOCK_N + c * 32 + (lane & 7) * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = (lane >> 3) + 4 * i;
                        aux[PREFETCH ? c : 0][i] = (r < rows_ok && lcol + 4 <= p.N)
                                                        ? __ldg(reinterpret_cast<const uint2*>(rp + (long long)r * p.ldc + lcol))
                                                        : make_uint2(0u, 0u);
                    }
                }
            }
            constexpr bool STAGE_BIAS = EPI >= 0 && (EPI & EPI_BIAS) != 0;
            float* bs = bias_stage + (warp - 2) * BLOCK_N;
            if (STAGE_BIAS) {
                const float* bg = p.bias + col_base + t.n_tile * BLOCK_N;
                const int ncols = p.N - t.n_tile * BLOCK_N;
#pragma unroll
                for (int j = lane * 4; j < BLOCK_N; j += 128) {
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j + 4 <= ncols) b4 = __ldg(reinterpret_cast<const float4*>(bg + j));
                    else {
                        if (j < ncols) b4.x = bg[j];
                        if (j + 1 < ncols) b4.y = bg[j + 1];
                        if (j + 2 < ncols) b4.z = bg[j + 2];
                    }
                    *reinterpret_cast<float4*>(bs + j) = b4;
                }
                __syncwarp();
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
#pragma unroll(PREFETCH ? BLOCK_N / 32 : 1)
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N + c * 32, raw);
                tmem_ld_wait();
                const int col0 = t.n_tile * BLOCK_N + c * 32;
                if (col0 >= p.N) break;
                // transpose through shared memory: thread = row  ->  8 lanes per row, 4 columns each
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + (lane * EPI_PITCH + 4 * j) * 4),
                                 "r"(raw[4 * j]), "r"(raw[4 * j + 1]), "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3])
                                 : "memory");
                __syncwarp();
                float v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                                 : "r"(st + (((lane >> 3) + 4 * i) * EPI_PITCH + (lane & 7) * 4) * 4)
                                 : "memory");
                __syncwarp();
                const int lcol = col0 + (lane & 7) * 4;
                const int cols_ok = p.N - lcol;
                if (rows_ok > 0 && has_work) {
                    if (PREFETCH) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 lo = unpack_bf16x2(aux[PREFETCH ? c : 0][i].x);
                            const float2 hi = unpack_bf16x2(aux[PREFETCH ? c : 0][i].y);
                            v[4 * i] = v[4 * i] * p.alpha + lo.x;
                            v[4 * i + 1] = v[4 * i + 1] * p.alpha + lo.y;
                            v[4 * i + 2] = v[4 * i + 2] * p.alpha + hi.x;
                            v[4 * i + 3] = v[4 * i + 3] * p.alpha + hi.y;
                        }
                        GemmParams q0 = p;  // alpha already applied; plain store of the sum
                        q0.alpha = 1.0f;
                        epilogue_chunk<bf16, 0>(q0, v, row_off0, rows_ok, col_base + lcol, cols_ok, lane);
                    } else if (EPI >= 0) {
                        epilogue_chunk<bf16, EPI>(p, v, row_off0, rows_ok, col_base + lcol, cols_ok, lane,
                                                  STAGE_BIAS ? bs + c * 32 + (lane & 7) * 4 : nullptr);
                    } else if (p.out_atomic || p.out_accumulate) {
                        epilogue_chunk_f32_accum(p, v, row_off0, rows_ok, col_base + lcol, cols_ok, lane);
                    } else if (p.c_f32) {
                        epilogue_chunk<float, -1>(p, v, row_off0, rows_ok, col_base + lcol, cols_ok, lane);
                    } else {
                        epilogue_chunk<bf16, -1>(p, v, row_off0, rows_ok, col_base + lcol, cols_ok, lane);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            sym == nullptr) {
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// (inner, rows, batch) bf16 view with a (64, box_rows, 1) box and 128-byte swizzle
static int make_map(CUtensorMap* m, const a2v_operand& o, int box_rows, const char* name) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        a2v_set_error("cuTensorMapEncodeTiled not available from the driver");
        return A2V_ERR_CUDA;
    }
    A2V_REQUIRE(o.ptr != nullptr, "gemm: operand %s is NULL", name);
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0, "gemm: operand %s not 16-byte aligned", name);
    A2V_REQUIRE(o.dim0 > 0 && o.dim1 > 0 && o.dim2 > 0, "gemm: operand %s has an empty extent", name);
    A2V_REQUIRE(o.stride1 % 8 == 0 && (o.dim2 == 1 || o.stride2 % 8 == 0),
                "gemm: operand %s strides must be multiples of 8 elements (16 bytes)", name);
    cuuint64_t dims[3] = {(cuuint64_t)o.dim0, (cuuint64_t)o.dim1, (cuuint64_t)o.dim2};
    cuuint64_t strides[2] = {(cuuint64_t)o.stride1 * 2, (cuuint64_t)(o.dim2 == 1 ? o.stride1 * o.dim1 : o.stride2) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(o.ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        a2v_set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %lld,%lld,%lld strides %lld,%lld)",
                      name, (int)r, (long long)o.dim0, (long long)o.dim1, (long long)o.dim2, (long long)o.stride1,
                      (long long)o.stride2);
        return A2V_ERR_CUDA;
    }
    return A2V_OK;
}

template <int BLOCK_N, int MODE, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BLOCK_N, epi_is_wide<EPI>()>;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, MODE, EPI>;
    if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (size_t)S::TOTAL) != A2V_OK) return A2V_ERR_CUDA;
    int grid = p.num_tiles < a2v_num_sms() ? p.num_tiles : a2v_num_sms();
    kern<<<grid, epi_is_wide<EPI>() ? GEMM_THREADS_WIDE : GEMM_THREADS, S::TOTAL, st>>>(ta, tb, p);
    return a2v_check_launch("gemm_tcgen05_kernel");
}

int gemm2cta_try(const a2v_gemm_desc* d, cudaStream_t st);  // gemm2_sm100.cu

// CTA-pair kernel for the plain Linear shapes: on unless A2V_GEMM_2CTA=0
static bool gemm2cta_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("A2V_GEMM_2CTA");
        on = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return on == 1;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_gemm(const a2v_gemm_desc* d, a2v_stream_t stream) {
    A2V_REQUIRE(d != nullptr, "gemm: NULL descriptor");
    A2V_REQUIRE(d->mode == 0 || d->mode == 1, "gemm: mode must be 0 (NT) or 1 (TN)");
    A2V_REQUIRE(d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "gemm: block_n must be 64, 128 or 256");
    A2V_REQUIRE(d->M > 0 && d->N > 0, "gemm: empty output (M=%d N=%d)", d->M, d->N);
    A2V_REQUIRE(d->taps >= 1 && d->batch >= 1 && d->groups >= 1, "gemm: taps/batch/groups must be >= 1");
    A2V_REQUIRE(d->a_tap_wrap >= 0 && d->a_grow_div >= 0, "gemm: a_tap_wrap / a_grow_div must be >= 0");
    A2V_REQUIRE(d->c != nullptr, "gemm: C is NULL");
    A2V_REQUIRE(d->c_dtype == A2V_F32 || d->c_dtype == A2V_BF16, "gemm: bad c_dtype");
    A2V_REQUIRE(!(d->out_atomic || d->out_accumulate) || d->c_dtype == A2V_F32,
                "gemm: atomic/accumulating output requires fp32 C");
    A2V_REQUIRE(!(d->out_atomic || d->out_accumulate) ||
                    (d->bias == nullptr && d->act == 0 && d->preact == nullptr && d->residual == nullptr &&
                     d->dgelu_u == nullptr),
                "gemm: accumulating output supports no fused epilogue");
    const int vec = d->c_dtype == A2V_F32 ? 4 : 8;
    A2V_REQUIRE(d->ldc % vec == 0 && d->c_group_stride % vec == 0 && d->c_tap_stride % vec == 0,
                "gemm: ldc / group / tap strides must keep 16-byte alignment");
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(d->c) & 15) == 0, "gemm: C not 16-byte aligned");

    if (gemm2cta_enabled()) {
        const int rc2 = gemm2cta_try(d, reinterpret_cast<cudaStream_t>(stream));
        if (rc2 >= 0) return rc2;
    }
    A2V_REQUIRE(d->colsum == nullptr, "gemm: the fused column-sum epilogue exists for the CTA-pair Linear shapes only "
                                      "(bf16, N %% 256 == 0, K %% 64 == 0, M >= 512, A2V_GEMM_2CTA on)");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.mode = d->mode;
    p.M = d->M;
    p.N = d->N;
    p.taps = d->taps;
    p.batch = d->batch;
    p.groups = d->groups;
    p.m_tiles = ceil_div(d->M, BLOCK_M);
    p.n_tiles = ceil_div(d->N, d->block_n);
    p.a_group_stride = d->a_group_stride;
    p.a_row_off = d->a_row_off;
    p.a_tap_rows = d->a_tap_rows;
    p.a_tap_cols = d->a_tap_cols;
    p.a_tap_wrap = d->a_tap_wrap;
    p.a_grow_add = d->a_grow_add;
    p.a_grow_div = d->a_grow_div;
    p.a_tap_col_stride = d->a_tap_col_stride > 0 ? d->a_tap_col_stride : d->a_tap_cols;
    p.b_group_stride = d->b_group_stride;
    p.b_row_off = d->b_row_off;
    p.b_tap_rows = d->b_tap_rows;
    p.c = d->c;
    p.c_f32 = d->c_dtype == A2V_F32;
    p.out_atomic = d->out_atomic;
    p.out_accumulate = d->out_accumulate;
    p.ldc = d->ldc;
    p.c_batch_stride = d->c_batch_stride;
    p.c_row_off = d->c_row_off;
    p.c_group_stride = d->c_group_stride;
    p.c_tap_stride = d->c_tap_stride;
    p.alpha = d->alpha;
    p.bias = d->bias;
    p.act = d->act;
    p.preact = d->preact;
    p.residual = d->residual;
    p.dgelu_u = d->dgelu_u;

    CUtensorMap ta, tb;
    int rc;
    if (d->mode == 0) {
        A2V_REQUIRE(d->k_per_tap > 0 && d->k_per_tap % 8 == 0, "gemm: k_per_tap must be a positive multiple of 8");
        A2V_REQUIRE(d->taps == 1 || d->k_per_tap % BLOCK_K == 0,
                    "gemm: multi-tap products need k_per_tap to be a multiple of 64");
        p.kb_per_tap = ceil_div(d->k_per_tap, BLOCK_K);
        p.k_splits = 1;
        p.num_tiles = p.n_tiles * p.m_tiles * p.batch * p.groups;
        if ((rc = make_map(&ta, d->a, BLOCK_M, "A")) != A2V_OK) return rc;
        if ((rc = make_map(&tb, d->b, d->block_n, "B")) != A2V_OK) return rc;
        A2V_REQUIRE(d->a_tap_cols == 0 || d->a_tap_wrap > 0, "gemm: NT mode takes a_tap_cols only with a_tap_wrap (strided conv)");
        A2V_REQUIRE(d->a_tap_wrap == 0 || (d->a_tap_cols > 0 && d->a_tap_cols % 8 == 0 && d->a_tap_col_stride % 8 == 0),
                    "gemm: wrapped tap addressing needs a_tap_cols / a_tap_col_stride in multiples of 8 columns");
    } else {
        A2V_REQUIRE(d->red_rows > 0 && d->k_splits >= 1, "gemm: TN mode needs red_rows > 0 and k_splits >= 1");
        A2V_REQUIRE(d->k_splits == 1 || d->out_atomic, "gemm: split-K requires out_atomic");
        A2V_REQUIRE(d->bias == nullptr, "gemm: TN mode has no bias epilogue");
        A2V_REQUIRE(d->a_tap_cols == 0 || (d->a_tap_cols > 0 && d->a_tap_cols % 64 == 0 && d->taps == 1),
                    "gemm: a_tap_cols must be a multiple of 64 (and taps == 1: the taps live in M)");
        A2V_REQUIRE(d->a_tap_wrap == 0 || d->a_tap_cols > 0, "gemm: a_tap_wrap needs a_tap_cols");
        A2V_REQUIRE(d->a_grow_div == 0, "gemm: a_grow_div is an NT-mode field");
        p.red_rows = d->red_rows;
        p.kb_per_batch = ceil_div(d->red_rows, BLOCK_K);
        p.k_splits = d->k_splits;
        p.kb_per_tap = 1;
        p.num_tiles = p.n_tiles * p.m_tiles * p.taps * p.groups * p.k_splits;
        if ((rc = make_map(&ta, d->a, 64, "A")) != A2V_OK) return rc;
        if ((rc = make_map(&tb, d->b, 64, "B")) != A2V_OK) return rc;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // hot bf16-output epilogues get straight-line code; everything else takes the run-time path
    int epi = -1;
    if (d->mode == 0 && d->c_dtype == A2V_BF16 && !d->out_atomic && !d->out_accumulate) {
        const int m = (d->bias ? EPI_BIAS : 0) | (d->preact ? EPI_PREACT : 0) | (d->act == 1 ? EPI_GELU : 0) |
                      (d->dgelu_u ? EPI_DGELU : 0) | (d->residual ? EPI_RES : 0);
        if (m == 0 || m == EPI_BIAS || m == (EPI_BIAS | EPI_GELU | EPI_PREACT) || m == (EPI_BIAS | EPI_GELU) ||
            m == EPI_DGELU || m == EPI_RES)
            epi = m;
        // the 8-warp epilogue moves whole 16-byte column groups
        if (epi >= 0 && (epi & EPI_DGELU) == 0 &&
            (d->N % 32 != 0 || d->ldc % 8 != 0 || d->c_group_stride % 8 != 0 ||
             (reinterpret_cast<uintptr_t>(d->c) & 15) != 0 || (reinterpret_cast<uintptr_t>(d->preact) & 15) != 0 ||
             (reinterpret_cast<uintptr_t>(d->residual) & 15) != 0 ||
             (reinterpret_cast<uintptr_t>(d->bias) & 15) != 0))
            epi = -1;
    }
    // staged bias vectors are read with 16-byte loads
    if (epi >= 0 && (epi & EPI_BIAS) && ((reinterpret_cast<uintptr_t>(d->bias) & 15) != 0 || d->c_group_stride % 4 != 0))
        epi = -1;
#define A2V_DISPATCH(BN)                                                                              \
    (d->mode == 1 ? launch_gemm<BN, 1, -1>(ta, tb, p, st)                                             \
     : epi == 0 ? launch_gemm<BN, 0, 0>(ta, tb, p, st)                                                \
     : epi == EPI_BIAS ? launch_gemm<BN, 0, EPI_BIAS>(ta, tb, p, st)                                  \
     : epi == (EPI_BIAS | EPI_GELU | EPI_PREACT) ? launch_gemm<BN, 0, (EPI_BIAS | EPI_GELU | EPI_PREACT)>(ta, tb, p, st) \
     : epi == (EPI_BIAS | EPI_GELU) ? launch_gemm<BN, 0, (EPI_BIAS | EPI_GELU)>(ta, tb, p, st)        \
     : epi == EPI_DGELU ? launch_gemm<BN, 0, EPI_DGELU>(ta, tb, p, st)                                \
     : epi == EPI_RES ? launch_gemm<BN, 0, EPI_RES>(ta, tb, p, st)                                    \
     : launch_gemm<BN, 0, -1>(ta, tb, p, st))
    if (d->block_n == 64) return A2V_DISPATCH(64);
    if (d->block_n == 128) return A2V_DISPATCH(128);
    return A2V_DISPATCH(256);
#undef A2V_DISPATCH
}
