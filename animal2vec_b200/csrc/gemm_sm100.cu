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
#include <string.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
    int mode;
    int M, N;
    int kb_per_tap, taps, batch, groups;
    int m_tiles, n_tiles;
    int a_group_stride, a_row_off, a_tap_rows;
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

template <int BLOCK_N>
struct GemmSmem {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
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

// Epilogue for one 32-column chunk of one accumulator row.
template <typename TC>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&v)[32], long long row_off, int gcol0,
                                               int ncols_valid) {
    TC* c = reinterpret_cast<TC*>(p.c) + row_off + gcol0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
    if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols_valid) v[i] += __ldg(p.bias + gcol0 + i);
    }
    const bool full = (ncols_valid == 32);
    if (p.preact != nullptr) {
        TC* u = reinterpret_cast<TC*>(p.preact) + row_off + gcol0;
        if (full) {
            store32(u, v);
        } else {
            for (int i = 0; i < ncols_valid; ++i) u[i] = from_f32<TC>(v[i]);
        }
    }
    if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_exact(v[i]);
    }
    if (p.dgelu_u != nullptr) {
        const TC* u = reinterpret_cast<const TC*>(p.dgelu_u) + row_off + gcol0;
        float t[32];
        if (full) {
            load32(u, t);
        } else {
            for (int i = 0; i < 32; ++i) t[i] = (i < ncols_valid) ? to_f32(u[i]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= gelu_exact_grad(t[i]);
    }
    if (p.residual != nullptr) {
        const TC* r = reinterpret_cast<const TC*>(p.residual) + row_off + gcol0;
        float t[32];
        if (full) {
            load32(r, t);
        } else {
            for (int i = 0; i < 32; ++i) t[i] = (i < ncols_valid) ? to_f32(r[i]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += t[i];
    }
    if (full) {
        store32(c, v);
    } else {
        for (int i = 0; i < ncols_valid; ++i) c[i] = from_f32<TC>(v[i]);
    }
}

__device__ __forceinline__ void epilogue_chunk_f32_accum(const GemmParams& p, float (&v)[32], long long row_off,
                                                         int gcol0, int ncols_valid) {
    float* c = reinterpret_cast<float*>(p.c) + row_off + gcol0;
    if (p.out_atomic) {
        for (int i = 0; i < ncols_valid; ++i) atomicAdd(c + i, v[i] * p.alpha);
    } else {
        for (int i = 0; i < ncols_valid; ++i) c[i] += v[i] * p.alpha;
    }
}

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
    using S = GemmSmem<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* tfull_bar = empty_bar + S::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

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
            mbar_init(&tempty_bar[i], 4);
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
        if (lane == 0) {
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
                        tma_load_3d(sa, &tmA, &full_bar[stage], t.g * p.a_group_stride + kc * BLOCK_K,
                                    t.m_tile * BLOCK_M + tap * p.a_tap_rows + p.a_row_off, t.b);
                        tma_load_3d(sb, &tmB, &full_bar[stage], kb * BLOCK_K,
                                    t.g * p.b_group_stride + t.n_tile * BLOCK_N, 0);
                    } else {
                        const int b = kb / p.kb_per_batch;
                        const int r0 = (kb - b * p.kb_per_batch) * BLOCK_K;
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i)
                            tma_load_3d(sa + i * (64 * BLOCK_K * 2), &tmA, &full_bar[stage],
                                        t.g * p.a_group_stride + t.m_tile * BLOCK_M + i * 64, r0, b);
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
        if (lane == 0) {
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
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            int kb0, kb1;
            tile_k_range(p, t, kb0, kb1);
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int row_in_block = t.m_tile * BLOCK_M + r;
            const bool row_ok = row_in_block < p.M;
            long long row_g;
            int col_base;
            if (MODE == 0) {
                row_g = (long long)t.b * p.c_batch_stride + p.c_row_off + row_in_block;
                col_base = t.g * p.c_group_stride;
            } else {
                row_g = (long long)t.g * p.c_group_stride + row_in_block;
                col_base = t.tap * p.c_tap_stride;
            }
            const long long row_off = row_g * p.ldc;
            const bool has_work = (kb1 > kb0);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N + c * 32, raw);
                tmem_ld_wait();
                const int col0 = t.n_tile * BLOCK_N + c * 32;
                int nvalid = p.N - col0;
                nvalid = nvalid > 32 ? 32 : nvalid;
                if (row_ok && nvalid > 0 && has_work) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                    if (p.out_atomic || p.out_accumulate) {
                        epilogue_chunk_f32_accum(p, v, row_off, col_base + col0, nvalid);
                    } else if (p.c_f32) {
                        epilogue_chunk<float>(p, v, row_off, col_base + col0, nvalid);
                    } else {
                        epilogue_chunk<bf16>(p, v, row_off, col_base + col0, nvalid);
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

template <int BLOCK_N, int MODE>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BLOCK_N>;
    static bool configured = false;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, MODE>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            a2v_set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return A2V_ERR_CUDA;
        }
        configured = true;
    }
    int grid = p.num_tiles < a2v_num_sms() ? p.num_tiles : a2v_num_sms();
    kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(ta, tb, p);
    return a2v_check_launch("gemm_tcgen05_kernel");
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_gemm(const a2v_gemm_desc* d, a2v_stream_t stream) {
    A2V_REQUIRE(d != nullptr, "gemm: NULL descriptor");
    A2V_REQUIRE(d->mode == 0 || d->mode == 1, "gemm: mode must be 0 (NT) or 1 (TN)");
    A2V_REQUIRE(d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "gemm: block_n must be 64, 128 or 256");
    A2V_REQUIRE(d->M > 0 && d->N > 0, "gemm: empty output (M=%d N=%d)", d->M, d->N);
    A2V_REQUIRE(d->taps >= 1 && d->batch >= 1 && d->groups >= 1, "gemm: taps/batch/groups must be >= 1");
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
    } else {
        A2V_REQUIRE(d->red_rows > 0 && d->k_splits >= 1, "gemm: TN mode needs red_rows > 0 and k_splits >= 1");
        A2V_REQUIRE(d->k_splits == 1 || d->out_atomic, "gemm: split-K requires out_atomic");
        A2V_REQUIRE(d->bias == nullptr, "gemm: TN mode has no bias epilogue");
        p.red_rows = d->red_rows;
        p.kb_per_batch = ceil_div(d->red_rows, BLOCK_K);
        p.k_splits = d->k_splits;
        p.kb_per_tap = 1;
        p.num_tiles = p.n_tiles * p.m_tiles * p.taps * p.groups * p.k_splits;
        if ((rc = make_map(&ta, d->a, 64, "A")) != A2V_OK) return rc;
        if ((rc = make_map(&tb, d->b, 64, "B")) != A2V_OK) return rc;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define A2V_DISPATCH(BN)                                                         \
    (d->mode == 0 ? launch_gemm<BN, 0>(ta, tb, p, st) : launch_gemm<BN, 1>(ta, tb, p, st))
    if (d->block_n == 64) return A2V_DISPATCH(64);
    if (d->block_n == 128) return A2V_DISPATCH(128);
    return A2V_DISPATCH(256);
#undef A2V_DISPATCH
}
