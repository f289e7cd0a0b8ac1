// Fused row-wise LayerNorm family (memory bound; one warp per row, 8/16-byte vector access).
//
//   z = a + dropout_b(b)                       (b optional)
//   n = (z - mean(z)) * rstd(z) * gamma + beta (gamma/beta optional)
//   y = dropout_out(act(n)) + post             (act: none | exact GELU | PSwish; post optional)
//
// covers, with channels-last activations:
//   Fp32LayerNorm(127)+PSwish and Fp32LayerNorm(512)+GELU of the feature extractor
//     (reference nn/utils.py:1105-1117, PSwish 1413-1435),
//   Fp32LayerNorm(512) of project_features (nn/modalities/audio.py:86),
//   LayerNorm(no affine)+GELU of the positional conv stack (audio.py:104-108) and of the
//     decoder blocks incl. the decoder residual (nn/modalities/modules.py:150-157,124-134),
//   the post-LN residual norms of AltBlock: LN(x + drop(attn)), LN(r + drop(mlp))
//     (modules.py:329-333) and BlockEncoder's LN -> dropout (modules.py:84-87).
// "Group-padded" channel layouts (gw stored channels per group of which gr are real) let
// 127- and 48-wide channel groups sit in 128/64-wide, TMA-friendly rows; pads stay zero.
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

struct RowLnParams {
    const void* a;
    const void* b;
    const float* gamma;
    const float* beta;
    const float* act_alpha;
    const float* act_beta;
    const void* post;
    void* y;
    float* mean;
    float* rstd;
    long long rows;
    int C, gw, gr;
    float eps;
    int act;
    float drop_b;
    unsigned long long seed_b;
    float drop_out;
    unsigned long long seed_out;
    // backward
    const void* dy;
    void* da;
    void* db;
    float* dgamma;
    float* dbeta;
    float* dact_alpha;
    float* dact_beta;
};

__device__ __forceinline__ int param_index(int c, int gw, int gr) { return (c / gw) * gr + (c % gw); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------
// Per-warp asynchronous row pipeline. Every warp owns a ring of STAGES shared-memory slots; lane 0
// fills a slot with up to three 1-D bulk copies (cp.async.bulk, completion on the slot's mbarrier),
// the warp copies the landed row into registers, immediately re-arms the slot with the row STAGES
// iterations ahead and only then does the arithmetic. Bytes in flight per SM are set by the ring
// size (tens of KB), not by register-limited occupancy, which is what a warp-per-row LayerNorm
// needs to reach HBM bandwidth.
// ---------------------------------------------------------------------------------------------
constexpr int ROWLN_WARPS = 8;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct RowPipe {
    unsigned char* slots;  // this warp's ring: stages x ntens x row_bytes
    uint64_t* bars;        // this warp's mbarriers [stages]
    int stages, row_bytes, ntens;
    const unsigned char* src[3];
    long long rows, nwarps;

    __device__ __forceinline__ void issue(int slot, long long row, int lane) {
        if (lane == 0 && row < rows) {
            fence_proxy_async();
            mbar_expect_tx(&bars[slot], (uint32_t)(ntens * row_bytes));
            unsigned char* dst = slots + (size_t)slot * ntens * row_bytes;
            for (int t = 0; t < ntens; ++t)
                bulk_load_1d(dst + (size_t)t * row_bytes, src[t] + row * row_bytes, (uint32_t)row_bytes, &bars[slot]);
        }
    }
    __device__ __forceinline__ const unsigned char* wait(int slot, int it) {
        mbar_wait(&bars[slot], (uint32_t)((it / stages) & 1));
        return slots + (size_t)slot * ntens * row_bytes;
    }
};

// shared layout: [params: nparam x C floats][rings: WARPS x stages x ntens x row_bytes][mbarriers]
template <typename T>
__device__ __forceinline__ RowPipe make_pipe(unsigned char* smem, int nparam_floats, int C, int stages, int ntens,
                                             const void* s0, const void* s1, const void* s2, long long rows,
                                             long long nwarps, long long warp0, int lane) {
    RowPipe rp;
    const int warp = threadIdx.x >> 5;
    rp.stages = stages;
    rp.ntens = ntens;
    rp.row_bytes = C * (int)sizeof(T);
    unsigned char* rings = smem + (size_t)nparam_floats * sizeof(float);
    rp.slots = rings + (size_t)warp * stages * ntens * rp.row_bytes;
    rp.bars = reinterpret_cast<uint64_t*>(rings + (size_t)ROWLN_WARPS * stages * ntens * rp.row_bytes) + warp * stages;
    rp.src[0] = reinterpret_cast<const unsigned char*>(s0);
    rp.src[1] = reinterpret_cast<const unsigned char*>(s1);
    rp.src[2] = reinterpret_cast<const unsigned char*>(s2);
    rp.rows = rows;
    rp.nwarps = nwarps;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&rp.bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    for (int s = 0; s < stages; ++s) rp.issue(s, warp0 + (long long)s * nwarps, lane);
    return rp;
}

// stored-channel-indexed parameter tables in shared memory (pads -> 0): no index math in the row loop
__device__ __forceinline__ void fill_param_table(float* dst, const float* src, int C, int gw, int gr, bool padded) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const bool real = !padded || (c % gw) < gr;
        dst[c] = (src != nullptr && real) ? src[padded ? param_index(c, gw, gr) : c] : 0.f;
    }
}

template <typename T, int NCH, bool FULL, int ACT, bool AFFINE>
__global__ void __launch_bounds__(256) rowln_fwd_kernel(const RowLnParams p, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    const int C = p.C;
    const bool padded = p.gr < p.gw;
    const int creal = (C / p.gw) * p.gr;
    const float inv_c = 1.0f / (float)creal;
    const float keep_b = p.drop_b > 0.f ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    const float keep_o = p.drop_out > 0.f ? 1.0f / (1.0f - p.drop_out) : 1.0f;
    T* Y = reinterpret_cast<T*>(p.y);
    const bool has_b = p.b != nullptr, has_post = p.post != nullptr;
    constexpr bool affine = AFFINE;

    float* tab = reinterpret_cast<float*>(smem_raw);  // [gamma | beta | act_alpha | act_beta] x C
    fill_param_table(tab, p.gamma, C, p.gw, p.gr, padded);
    fill_param_table(tab + C, p.beta, C, p.gw, p.gr, padded);
    fill_param_table(tab + 2 * C, p.act_alpha, C, p.gw, p.gr, padded);
    fill_param_table(tab + 3 * C, p.act_beta, C, p.gw, p.gr, padded);
    // per-lane pad masks (bit j of nibble i): depends on the channel only, hoisted out of the row loop
    unsigned realmask = 0;  // FULL: every stored channel of every chunk is real, the tests fold away
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (lane + 32 * i) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (FULL || (c + j < C && (!padded || ((c + j) % p.gw) < p.gr))) realmask |= 1u << (4 * i + j);
    }
#define RL_REAL(i, j) (FULL || ((realmask >> (4 * (i) + (j))) & 1u))
    __syncthreads();

    const int ntens = 1 + (has_b ? 1 : 0) + (has_post ? 1 : 0);
    RowPipe rp = make_pipe<T>(smem_raw, 4 * C, C, stages, ntens, p.a, has_b ? p.b : p.post, p.post, p.rows, nwarps,
                              warp0, lane);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        const unsigned char* sm = rp.wait(slot, it);
        const T* sa = reinterpret_cast<const T*>(sm);
        const T* sb = reinterpret_cast<const T*>(sm + rp.row_bytes);
        const T* sp = reinterpret_cast<const T*>(sm + (size_t)(has_b ? 2 : 1) * rp.row_bytes);
        float z[NCH][4], post[NCH][4];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) z[i][j] = post[i][j] = 0.f;
            if (FULL || c < C) {
                load4(sa + c, z[i]);
                if (has_b) {
                    float t[4];
                    load4(sb + c, t);
                    if (p.drop_b > 0.f) {
                        bool k[4];
                        drop_keep4(p.seed_b, (unsigned long long)(row * C + c) >> 2, p.drop_b, k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = k[j] ? t[j] * keep_b : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[i][j] += t[j];
                }
                if (has_post) load4(sp + c, post[i]);
            }
        }
        __syncwarp();
        rp.issue(slot, row + (long long)stages * nwarps, lane);  // slot is in registers now: re-arm it

        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!RL_REAL(i, j)) z[i][j] = 0.f;
                s += z[i][j];
            }
        const float mean = warp_sum(s) * inv_c;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = z[i][j] - mean;
                v += RL_REAL(i, j) ? d * d : 0.f;
            }
        const float rstd = rsqrtf(warp_sum(v) * inv_c + p.eps);
        if (lane == 0) {
            if (p.mean != nullptr) p.mean[row] = mean;
            if (p.rstd != nullptr) p.rstd[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (FULL || c < C) {
                float o[4];
                bool ko[4] = {true, true, true, true};
                if (p.drop_out > 0.f) drop_keep4(p.seed_out, (unsigned long long)(row * C + c) >> 2, p.drop_out, ko);
                float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
                if (affine) {
                    load4(tab + c, ga);
                    load4(tab + C + c, be);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = RL_REAL(i, j);
                    float n = (z[i][j] - mean) * rstd;
                    if (affine) n = n * ga[j] + be[j];
                    float y;
                    if (ACT == 1) {
                        y = gelu_t<T>(n);
                    } else if (ACT == 2) {
                        y = n * tab[2 * C + c + j] * sigmoidf_(tab[3 * C + c + j] * n);
                    } else {
                        y = n;
                    }
                    if (p.drop_out > 0.f) y = ko[j] ? y * keep_o : 0.f;
                    y += post[i][j];
                    o[j] = real ? y : 0.f;
                }
                store4(Y + row * C + c, o);
            }
        }
    }
}

template <typename T, int NCH, bool FULL, int ACT, bool AFFINE>
__global__ void __launch_bounds__(256) rowln_bwd_kernel(const RowLnParams p, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    const int C = p.C;
    const bool padded = p.gr < p.gw;
    const int creal = (C / p.gw) * p.gr;
    const float inv_c = 1.0f / (float)creal;
    const float keep_b = p.drop_b > 0.f ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    const float keep_o = p.drop_out > 0.f ? 1.0f / (1.0f - p.drop_out) : 1.0f;
    T* DA = reinterpret_cast<T*>(p.da);
    T* DB = reinterpret_cast<T*>(p.db);
    const bool has_b = p.b != nullptr;
    constexpr bool affine = AFFINE;
    const bool want_affine = AFFINE && p.dgamma != nullptr;
    const bool want_act = ACT == 2 && p.dact_alpha != nullptr;

    float* tab = reinterpret_cast<float*>(smem_raw);  // [gamma | beta | act_alpha | act_beta] x C
    float* sred = tab + 4 * C;                         // [dgamma | dbeta | dalpha | dabeta] x C partials
    fill_param_table(tab, p.gamma, C, p.gw, p.gr, padded);
    fill_param_table(tab + C, p.beta, C, p.gw, p.gr, padded);
    fill_param_table(tab + 2 * C, p.act_alpha, C, p.gw, p.gr, padded);
    fill_param_table(tab + 3 * C, p.act_beta, C, p.gw, p.gr, padded);
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) sred[i] = 0.f;
    unsigned realmask = 0;  // FULL: every stored channel of every chunk is real, the tests fold away
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (lane + 32 * i) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (FULL || (c + j < C && (!padded || ((c + j) % p.gw) < p.gr))) realmask |= 1u << (4 * i + j);
    }
#define RL_REAL(i, j) (FULL || ((realmask >> (4 * (i) + (j))) & 1u))
    __syncthreads();

    constexpr int NACC = AFFINE ? NCH : 1;
    float acc_g[NACC][4], acc_b[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_g[i][j] = acc_b[i][j] = 0.f;

    const int ntens = has_b ? 3 : 2;
    RowPipe rp = make_pipe<T>(smem_raw, 8 * C, C, stages, ntens, p.a, p.dy, p.b, p.rows, nwarps, warp0, lane);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        const unsigned char* sm = rp.wait(slot, it);
        const T* sa = reinterpret_cast<const T*>(sm);
        const T* sdy = reinterpret_cast<const T*>(sm + rp.row_bytes);
        const T* sb = reinterpret_cast<const T*>(sm + (size_t)2 * rp.row_bytes);
        float xh[NCH][4], g[NCH][4];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) xh[i][j] = g[i][j] = 0.f;
            if (FULL || c < C) {
                load4(sa + c, xh[i]);  // z for now
                if (has_b) {
                    float t[4];
                    load4(sb + c, t);
                    if (p.drop_b > 0.f) {
                        bool k[4];
                        drop_keep4(p.seed_b, (unsigned long long)(row * C + c) >> 2, p.drop_b, k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = k[j] ? t[j] * keep_b : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) xh[i][j] += t[j];
                }
                load4(sdy + c, g[i]);  // dy for now
            }
        }
        __syncwarp();
        rp.issue(slot, row + (long long)stages * nwarps, lane);

        const float mean = p.mean[row], rstd = p.rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (FULL || c < C) {
                if (p.drop_out > 0.f) {
                    bool ko[4];
                    drop_keep4(p.seed_out, (unsigned long long)(row * C + c) >> 2, p.drop_out, ko);
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[i][j] = ko[j] ? g[i][j] * keep_o : 0.f;
                }
                float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
                if (affine) {
                    load4(tab + c, ga);
                    load4(tab + C + c, be);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = RL_REAL(i, j);
                    const float dyv = g[i][j];
                    const float x = (xh[i][j] - mean) * rstd;
                    const float n = x * ga[j] + be[j];
                    float dn;
                    if (ACT == 1) {
                        dn = dyv * gelu_grad_t<T>(n);
                    } else if (ACT == 2) {
                        const float al = tab[2 * C + c + j], bt = tab[3 * C + c + j];
                        const float sg = sigmoidf_(bt * n);
                        dn = dyv * al * (sg + n * bt * sg * (1.f - sg));
                        if (want_act && real) {
                            atomicAdd(&sred[2 * C + c + j], dyv * n * sg);
                            atomicAdd(&sred[3 * C + c + j], dyv * al * n * n * sg * (1.f - sg));
                        }
                    } else {
                        dn = dyv;
                    }
                    if (!real) dn = 0.f;
                    if (AFFINE) {
                        acc_g[AFFINE ? i : 0][j] += dn * x;
                        acc_b[AFFINE ? i : 0][j] += dn;
                    }
                    xh[i][j] = real ? x : 0.f;
                    g[i][j] = dn * ga[j];
                    s1 += g[i][j];
                    s2 += g[i][j] * xh[i][j];
                }
            }
        }
        s1 = warp_sum(s1) * inv_c;
        s2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (FULL || c < C) {
                float dz[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = RL_REAL(i, j);
                    dz[j] = real ? rstd * (g[i][j] - s1 - xh[i][j] * s2) : 0.f;
                }
                if (DA != nullptr) store4(DA + row * C + c, dz);
                if (DB != nullptr) {
                    if (p.drop_b > 0.f) {
                        bool k[4];
                        drop_keep4(p.seed_b, (unsigned long long)(row * C + c) >> 2, p.drop_b, k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dz[j] = k[j] ? dz[j] * keep_b : 0.f;
                    }
                    store4(DB + row * C + c, dz);
                }
            }
        }
    }
    if (want_affine) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (FULL || c < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&sred[c + j], acc_g[AFFINE ? i : 0][j]);
                    atomicAdd(&sred[C + c + j], acc_b[AFFINE ? i : 0][j]);
                }
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const bool real = !padded || (c % p.gw) < p.gr;
        if (!real) continue;
        const int pi = padded ? param_index(c, p.gw, p.gr) : c;
        if (want_affine) {
            atomicAdd(p.dgamma + pi, sred[c]);
            if (p.dbeta != nullptr) atomicAdd(p.dbeta + pi, sred[C + c]);
        }
        if (want_act) {
            atomicAdd(p.dact_alpha + pi, sred[2 * C + c]);
            atomicAdd(p.dact_beta + pi, sred[3 * C + c]);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Specialised bf16 kernels for the two row shapes that dominate the step: LayerNorm(no affine) + GELU
// over the student's positional-conv rows (1024 contiguous channels) and over the group-padded decoder
// rows (16 groups x 48 real of 64 stored channels, optional residual). Compared with the generic kernels
// above everything is resolved at compile time (no run-time feature tests inside the unrolled loops, a
// fifth of the SASS, under 128 registers so two CTAs share an SM), every lane owns 16-byte chunks of
// REAL channels only (pad chunks cost no arithmetic) and GELU costs 6 instructions (tanh form with one MUFU
// tanh.approx; both kernels are issue-bound, not HBM-bound, with the 14-instruction erfc form).
// ---------------------------------------------------------------------------------------------
struct ChunkMap {
    int cs, cr, ngroups;  // stored / real 16-byte chunks per channel group, groups per row
    __device__ __forceinline__ int stored_chunk(int r) const { return (r / cr) * cs + (r % cr); }
};

// ring slot fill: ONE bulk copy per tensor row (pads included -- one copy per channel group of the real bytes
// only was measured 45% slower: the copy issue rate, not the bytes, limits a 2 KB row)
template <int NTENS>
__device__ __forceinline__ void fast_issue(unsigned char* slot, uint64_t* bar, const unsigned char* const* src,
                                           long long row, long long rows, int row_bytes, int lane) {
    if (lane == 0 && row < rows) {
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(NTENS * row_bytes));
#pragma unroll
        for (int t = 0; t < NTENS; ++t)
            bulk_load_1d(slot + (size_t)t * row_bytes, src[t] + row * row_bytes, (uint32_t)row_bytes, bar);
    }
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    return v;
}

constexpr int FAST_STAGES_MAX = 8;

// forward: y = GELU(LN(a)) (+ post); NV = 16-byte real chunks per lane (32 * NV * 8 real channels per row)
template <int NV, bool HAS_POST>
__global__ void __launch_bounds__(256, 2) rowln_gelu_fwd_kernel(const RowLnParams p, const int stages, const ChunkMap cm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NTENS = HAS_POST ? 2 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    const int row_bytes = p.C * 2;
    unsigned char* ring = smem_raw + (size_t)warp * stages * NTENS * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ROWLN_WARPS * stages * NTENS * row_bytes) + warp * FAST_STAGES_MAX;
    const unsigned char* src[NTENS];
    src[0] = reinterpret_cast<const unsigned char*>(p.a);
    if (HAS_POST) src[NTENS - 1] = reinterpret_cast<const unsigned char*>(p.post);
    int off[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) off[i] = cm.stored_chunk(lane + 32 * i) * 16;
    const int npad = (cm.cs - cm.cr) * cm.ngroups;
    const float inv_c = 1.0f / (float)(NV * 256);
    unsigned char* Y = reinterpret_cast<unsigned char*>(p.y);
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    for (int s = 0; s < stages; ++s)
        fast_issue<NTENS>(ring + (size_t)s * NTENS * row_bytes, &bars[s], src, warp0 + (long long)s * nwarps, p.rows, row_bytes, lane);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        mbar_wait(&bars[slot], (uint32_t)((it / stages) & 1));
        const unsigned char* sm = ring + (size_t)slot * NTENS * row_bytes;
        uint4 va[NV], vp[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            va[i] = *reinterpret_cast<const uint4*>(sm + off[i]);
            if (HAS_POST) vp[i] = *reinterpret_cast<const uint4*>(sm + row_bytes + off[i]);
        }
        __syncwarp();
        fast_issue<NTENS>(ring + (size_t)slot * NTENS * row_bytes, &bars[slot], src, row + (long long)stages * nwarps, p.rows,
                          row_bytes, lane);
        float z[NV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            unpack8(va[i], z[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += z[i][j];
        }
        const float mean = warp_sum(s) * inv_c;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                z[i][j] -= mean;
                v = fmaf(z[i][j], z[i][j], v);
            }
        const float rstd = rsqrtf(warp_sum(v) * inv_c + p.eps);
        if (lane == 0) {
            if (p.mean != nullptr) p.mean[row] = mean;
            if (p.rstd != nullptr) p.rstd[row] = rstd;
        }
        unsigned char* yrow = Y + row * row_bytes;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8], po[8];
            if (HAS_POST) unpack8(vp[i], po);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = gelu_tanh_fast(z[i][j] * rstd);
                if (HAS_POST) o[j] += po[j];
            }
            *reinterpret_cast<uint4*>(yrow + off[i]) = pack8(o);
        }
        for (int q = lane; q < npad; q += 32) {  // pad chunks stay exact zeros
            const int g = q / (cm.cs - cm.cr), k = q % (cm.cs - cm.cr);
            *reinterpret_cast<uint4*>(yrow + (size_t)(g * cm.cs + cm.cr + k) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// backward: da = dLN(dy * GELU'(n)) from the saved row statistics
template <int NV>
__global__ void __launch_bounds__(256, 2) rowln_gelu_bwd_kernel(const RowLnParams p, const int stages, const ChunkMap cm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NTENS = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    const int row_bytes = p.C * 2;
    unsigned char* ring = smem_raw + (size_t)warp * stages * NTENS * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ROWLN_WARPS * stages * NTENS * row_bytes) + warp * FAST_STAGES_MAX;
    const unsigned char* src[2] = {reinterpret_cast<const unsigned char*>(p.a), reinterpret_cast<const unsigned char*>(p.dy)};
    int off[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) off[i] = cm.stored_chunk(lane + 32 * i) * 16;
    const int npad = (cm.cs - cm.cr) * cm.ngroups;
    const float inv_c = 1.0f / (float)(NV * 256);
    unsigned char* DA = reinterpret_cast<unsigned char*>(p.da);
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    for (int s = 0; s < stages; ++s)
        fast_issue<NTENS>(ring + (size_t)s * NTENS * row_bytes, &bars[s], src, warp0 + (long long)s * nwarps, p.rows, row_bytes, lane);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        const float mean = p.mean[row], rstd = p.rstd[row];
        mbar_wait(&bars[slot], (uint32_t)((it / stages) & 1));
        const unsigned char* sm = ring + (size_t)slot * NTENS * row_bytes;
        uint4 va[NV], vg[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            va[i] = *reinterpret_cast<const uint4*>(sm + off[i]);
            vg[i] = *reinterpret_cast<const uint4*>(sm + row_bytes + off[i]);
        }
        __syncwarp();
        fast_issue<NTENS>(ring + (size_t)slot * NTENS * row_bytes, &bars[slot], src, row + (long long)stages * nwarps, p.rows,
                          row_bytes, lane);
        float xh[NV][8], g[NV][8];
        float s1 = 0.f, s2 = 0.f;
        const float nmr = -mean * rstd;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            unpack8(va[i], xh[i]);
            unpack8(vg[i], g[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[i][j] = fmaf(xh[i][j], rstd, nmr);
                g[i][j] *= gelu_tanh_fast_grad(xh[i][j]);
                s1 += g[i][j];
                s2 = fmaf(g[i][j], xh[i][j], s2);
            }
        }
        s1 = warp_sum(s1) * inv_c;
        s2 = warp_sum(s2) * inv_c;
        unsigned char* drow = DA + row * row_bytes;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
            *reinterpret_cast<uint4*>(drow + off[i]) = pack8(o);
        }
        for (int q = lane; q < npad; q += 32) {
            const int gq = q / (cm.cs - cm.cr), k = q % (cm.cs - cm.cr);
            *reinterpret_cast<uint4*>(drow + (size_t)(gq * cm.cs + cm.cr + k) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Feature-extractor layer 0: Fp32LayerNorm(127) + PSwish over (B * 80000) rows of 128 stored channels
// (reference nn/utils.py:1107-1117,1413-1435). A 256-byte row is too short for a warp: here a warp owns
// FOUR consecutive rows (1 KB, two fully coalesced 16-byte loads per lane), sixteen lanes per row and
// eight channels per lane, so a lane keeps its 8 x 4 parameters (and, backward, its 8 x 4 parameter
// gradient accumulators) in registers for the whole kernel and the row reductions take four shuffles.
// The next four rows are prefetched into registers while the current ones are processed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float half16_sum(float v) {  // sum over the 16 lanes that share a row
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Row128Lane {
    float ga[8], be[8], al[8], nb[8];  // gamma, beta, PSwish alpha, -log2(e) * PSwish beta (pads: 0)
};

__device__ __forceinline__ Row128Lane load_row128_params(const RowLnParams& p, int ch0) {
    Row128Lane q;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch0 + j;
        const bool real = c < p.gr;
        q.ga[j] = real ? p.gamma[c] : 0.f;
        q.be[j] = real ? p.beta[c] : 0.f;
        q.al[j] = real ? p.act_alpha[c] : 0.f;
        q.nb[j] = real ? -1.4426950408889634f * p.act_beta[c] : 0.f;
    }
    return q;
}

__global__ void __launch_bounds__(512, 1) rowln128_pswish_fwd_kernel(const RowLnParams p) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const int ch0 = (lane & 15) * 8;
    const int padj = p.gr - ch0;  // elements j >= padj of this lane are padding (>= 8: none)
    const Row128Lane q = load_row128_params(p, ch0);
    const float inv_c = 1.0f / (float)p.gr;
    const long long groups = (p.rows + 3) >> 2;  // four rows per warp iteration
    const uint4* A = reinterpret_cast<const uint4*>(p.a);
    uint4* Y = reinterpret_cast<uint4*>(p.y);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    auto fetch = [&](long long g, uint4 (&v)[2]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long row = g * 4 + 2 * h + (lane >> 4);
            v[h] = (g < groups && row < p.rows) ? __ldg(A + g * 64 + h * 32 + lane) : zero4;
        }
    };
    uint4 cur[2], nxt[2];
    fetch(gw, cur);
    for (long long g = gw; g < groups; g += nw) {
        fetch(g + nw, nxt);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long row = g * 4 + 2 * h + (lane >> 4);
            float z[8];
            unpack8(cur[h], z);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += z[j];
            const float mean = half16_sum(s) * inv_c;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                z[j] = j < padj ? z[j] - mean : 0.f;
                v = fmaf(z[j], z[j], v);
            }
            const float rstd = rsqrtf(half16_sum(v) * inv_c + p.eps);
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float n = fmaf(z[j] * rstd, q.ga[j], q.be[j]);
                const float sg = rcp_approx(1.0f + ex2_approx_f(q.nb[j] * n));  // sigmoid(beta * n)
                o[j] = (n * q.al[j]) * sg;
            }
            if (row < p.rows) {
                Y[g * 64 + h * 32 + lane] = pack8(o);
                if ((lane & 15) == 0) {
                    if (p.mean != nullptr) p.mean[row] = mean;
                    if (p.rstd != nullptr) p.rstd[row] = rstd;
                }
            }
        }
        cur[0] = nxt[0];
        cur[1] = nxt[1];
    }
}

__global__ void __launch_bounds__(384, 1) rowln128_pswish_bwd_kernel(const RowLnParams p) {
    __shared__ float sred[4][128];
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const int ch0 = (lane & 15) * 8;
    const int padj = p.gr - ch0;
    const Row128Lane q = load_row128_params(p, ch0);
    const float inv_c = 1.0f / (float)p.gr;
    const long long groups = (p.rows + 3) >> 2;
    const uint4* A = reinterpret_cast<const uint4*>(p.a);
    const uint4* DY = reinterpret_cast<const uint4*>(p.dy);
    uint4* DA = reinterpret_cast<uint4*>(p.da);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) (&sred[0][0])[i] = 0.f;
    __syncthreads();
    float acc_g[8], acc_b[8], acc_al[8], acc_ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[j] = acc_b[j] = acc_al[j] = acc_ab[j] = 0.f;
    auto fetch = [&](long long g, uint4 (&va)[2], uint4 (&vg)[2], float (&mu)[2], float (&rs)[2]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long row = g * 4 + 2 * h + (lane >> 4);
            const bool ok = g < groups && row < p.rows;
            va[h] = ok ? __ldg(A + g * 64 + h * 32 + lane) : zero4;
            vg[h] = ok ? __ldg(DY + g * 64 + h * 32 + lane) : zero4;
            mu[h] = ok ? __ldg(p.mean + row) : 0.f;
            rs[h] = ok ? __ldg(p.rstd + row) : 0.f;
        }
    };
    uint4 ca[2], cg[2], na[2], ng[2];
    float cm[2], cr[2], nm[2], nr[2];
    fetch(gw, ca, cg, cm, cr);
    for (long long g = gw; g < groups; g += nw) {
        fetch(g + nw, na, ng, nm, nr);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long row = g * 4 + 2 * h + (lane >> 4);
            float xh[8], gy[8];
            unpack8(ca[h], xh);
            unpack8(cg[h], gy);
            const float rstd = cr[h], nmr = -cm[h] * cr[h];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[j] = j < padj ? fmaf(xh[j], rstd, nmr) : 0.f;
                const float n = fmaf(xh[j], q.ga[j], q.be[j]);
                const float sg = rcp_approx(1.0f + ex2_approx_f(q.nb[j] * n));
                const float bt = q.nb[j] * -0.6931471805599453f;  // PSwish beta
                const float t = n * sg * (1.0f - sg);             // n * sigma'
                const float dyv = gy[j];
                acc_al[j] = fmaf(dyv, n * sg, acc_al[j]);
                acc_ab[j] = fmaf(dyv * q.al[j], n * t, acc_ab[j]);
                const float dn = dyv * q.al[j] * fmaf(bt, t, sg);
                acc_g[j] = fmaf(dn, xh[j], acc_g[j]);
                acc_b[j] += dn;
                gy[j] = dn * q.ga[j];
                s1 += gy[j];
                s2 = fmaf(gy[j], xh[j], s2);
            }
            s1 = half16_sum(s1) * inv_c;
            s2 = half16_sum(s2) * inv_c;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = j < padj ? rstd * (gy[j] - s1 - xh[j] * s2) : 0.f;
            if (row < p.rows) DA[g * 64 + h * 32 + lane] = pack8(o);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) { ca[h] = na[h]; cg[h] = ng[h]; cm[h] = nm[h]; cr[h] = nr[h]; }
    }
    // lanes l and l ^ 16 own the same channels
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc_g[j] += __shfl_xor_sync(0xffffffffu, acc_g[j], 16);
        acc_b[j] += __shfl_xor_sync(0xffffffffu, acc_b[j], 16);
        acc_al[j] += __shfl_xor_sync(0xffffffffu, acc_al[j], 16);
        acc_ab[j] += __shfl_xor_sync(0xffffffffu, acc_ab[j], 16);
    }
    if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sred[0][ch0 + j], acc_g[j]);
            atomicAdd(&sred[1][ch0 + j], acc_b[j]);
            atomicAdd(&sred[2][ch0 + j], acc_al[j]);
            atomicAdd(&sred[3][ch0 + j], acc_ab[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.gr; c += blockDim.x) {
        if (p.dgamma != nullptr) atomicAdd(p.dgamma + c, sred[0][c]);
        if (p.dbeta != nullptr) atomicAdd(p.dbeta + c, sred[1][c]);
        if (p.dact_alpha != nullptr) atomicAdd(p.dact_alpha + c, sred[2][c]);
        if (p.dact_beta != nullptr) atomicAdd(p.dact_beta + c, sred[3][c]);
    }
}

static int launch_rowln128(const RowLnParams& p, bool bwd, cudaStream_t st) {
    if (p.C != 128 || p.gw != 128 || p.gr < 121 || p.act != 2 || p.gamma == nullptr || p.beta == nullptr ||
        p.b != nullptr || p.post != nullptr || p.drop_out > 0.f)
        return -1;
    if ((reinterpret_cast<uintptr_t>(p.a) & 15) != 0) return -1;
    if (bwd && (p.da == nullptr || p.db != nullptr || (reinterpret_cast<uintptr_t>(p.dy) & 15) != 0)) return -1;
    const long long groups = (p.rows + 3) >> 2;
    const int threads = bwd ? 384 : 512;  // backward: 170 registers per thread for the parameter-gradient accumulators
    long long blocks = ceil_div64(groups, threads / 32);
    const long long cap = a2v_num_sms();
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (bwd)
        rowln128_pswish_bwd_kernel<<<grid, 384, 0, st>>>(p);
    else
        rowln128_pswish_fwd_kernel<<<grid, 512, 0, st>>>(p);
    return a2v_check_launch(bwd ? "rowln_bwd(c128)" : "rowln_fwd(c128)");
}

// ---------------------------------------------------------------------------------------------
// Specialised bf16 kernels for the post-LN residual norms of AltBlock (reference nn/modalities/modules.py:
// 329-333: x = LN1(x + drop(attn)), x = LN2(r + drop(mlp))) and BlockEncoder's first norm without a second
// operand: y = LN(a + dropout_b(b)) * gamma + beta over rows of NV*256 contiguous channels. 96 launches
// forward and 48 backward per step at the large config. Everything is resolved at compile time; a lane
// owns NV 16-byte chunks; the backward keeps gamma and the three column accumulators (dgamma, dbeta and
// the column sum of db = the bias gradient of the Linear that produced b) in registers for the whole
// kernel, and the dropout keep flags of a row in ONE 32-bit register (hashed once, used for z and for db).
// Dropout indexing is the generic kernels' (one 64-bit hash per 4 consecutive channels).
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ unsigned res_keep_bits(unsigned long long seed, long long row, int C, const int (&off)[NV],
                                                  float drop) {
    unsigned bits = 0;
    const uint32_t thr = (uint32_t)(drop * 65536.0f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const unsigned long long idx4 = (unsigned long long)(row * C + (off[i] >> 1)) >> 2;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const uint64_t h = rng64(seed, idx4 + hh);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((uint32_t)(h >> (16 * j)) & 0xffffu) >= thr) bits |= 1u << (8 * i + 4 * hh + j);
        }
    }
    return bits;
}

template <int NV, bool HAS_B, bool DROP>
__global__ void __launch_bounds__(256, 2) resln_fwd_kernel(const RowLnParams p, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NTENS = HAS_B ? 2 : 1;
    constexpr int C = NV * 256;
    constexpr int row_bytes = C * 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    float* tab = reinterpret_cast<float*>(smem_raw);  // [gamma | beta] x C
    unsigned char* rings = smem_raw + 2 * C * sizeof(float);
    unsigned char* ring = rings + (size_t)warp * stages * NTENS * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(rings + (size_t)ROWLN_WARPS * stages * NTENS * row_bytes) + warp * FAST_STAGES_MAX;
    const unsigned char* src[NTENS];
    src[0] = reinterpret_cast<const unsigned char*>(p.a);
    if (HAS_B) src[NTENS - 1] = reinterpret_cast<const unsigned char*>(p.b);
    int off[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) off[i] = (lane + 32 * i) * 16;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        tab[c] = p.gamma[c];
        tab[C + c] = p.beta != nullptr ? p.beta[c] : 0.f;
    }
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    for (int s = 0; s < stages; ++s)
        fast_issue<NTENS>(ring + (size_t)s * NTENS * row_bytes, &bars[s], src, warp0 + (long long)s * nwarps, p.rows, row_bytes, lane);
    const float inv_c = 1.0f / (float)C;
    const float keep_b = DROP ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    unsigned char* Y = reinterpret_cast<unsigned char*>(p.y);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        unsigned keep = 0xffffffffu;
        if (HAS_B && DROP) keep = res_keep_bits<NV>(p.seed_b, row, C, off, p.drop_b);  // before the wait: overlaps the copy
        mbar_wait(&bars[slot], (uint32_t)((it / stages) & 1));
        const unsigned char* sm = ring + (size_t)slot * NTENS * row_bytes;
        uint4 va[NV], vb[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            va[i] = *reinterpret_cast<const uint4*>(sm + off[i]);
            if (HAS_B) vb[i] = *reinterpret_cast<const uint4*>(sm + row_bytes + off[i]);
        }
        __syncwarp();
        fast_issue<NTENS>(ring + (size_t)slot * NTENS * row_bytes, &bars[slot], src, row + (long long)stages * nwarps, p.rows,
                          row_bytes, lane);
        float z[NV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            unpack8(va[i], z[i]);
            if (HAS_B) {
                float t[8];
                unpack8(vb[i], t);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (DROP) t[j] = ((keep >> (8 * i + j)) & 1u) ? t[j] * keep_b : 0.f;
                    z[i][j] += t[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s += z[i][j];
        }
        const float mean = warp_sum(s) * inv_c;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                z[i][j] -= mean;
                v = fmaf(z[i][j], z[i][j], v);
            }
        const float rstd = rsqrtf(warp_sum(v) * inv_c + p.eps);
        if (lane == 0) {
            if (p.mean != nullptr) p.mean[row] = mean;
            if (p.rstd != nullptr) p.rstd[row] = rstd;
        }
        unsigned char* yrow = Y + row * row_bytes;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8];
            const float4 g0 = *reinterpret_cast<const float4*>(tab + (off[i] >> 1));
            const float4 g1 = *reinterpret_cast<const float4*>(tab + (off[i] >> 1) + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(tab + C + (off[i] >> 1));
            const float4 b1 = *reinterpret_cast<const float4*>(tab + C + (off[i] >> 1) + 4);
            o[0] = fmaf(z[i][0] * rstd, g0.x, b0.x); o[1] = fmaf(z[i][1] * rstd, g0.y, b0.y);
            o[2] = fmaf(z[i][2] * rstd, g0.z, b0.z); o[3] = fmaf(z[i][3] * rstd, g0.w, b0.w);
            o[4] = fmaf(z[i][4] * rstd, g1.x, b1.x); o[5] = fmaf(z[i][5] * rstd, g1.y, b1.y);
            o[6] = fmaf(z[i][6] * rstd, g1.z, b1.z); o[7] = fmaf(z[i][7] * rstd, g1.w, b1.w);
            *reinterpret_cast<uint4*>(yrow + off[i]) = pack8(o);
        }
    }
}

// backward: da = dLN(dy * gamma); db = dropout_b-masked da; dgamma += sum dy*xhat, dbeta += sum dy, dbias_b += sum db
template <int NV, bool HAS_B, bool DROP>
__global__ void __launch_bounds__(256, 1) resln_bwd_kernel(const RowLnParams p, const int stages, float* dbias_b) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NTENS = HAS_B ? 3 : 2;
    constexpr int C = NV * 256;
    constexpr int row_bytes = C * 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * ROWLN_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * ROWLN_WARPS;
    unsigned char* ring = smem_raw + (size_t)warp * stages * NTENS * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ROWLN_WARPS * stages * NTENS * row_bytes) + warp * FAST_STAGES_MAX;
    const unsigned char* src[NTENS];
    src[0] = reinterpret_cast<const unsigned char*>(p.a);
    src[1] = reinterpret_cast<const unsigned char*>(p.dy);
    if (HAS_B) src[NTENS - 1] = reinterpret_cast<const unsigned char*>(p.b);
    int off[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) off[i] = (lane + 32 * i) * 16;
    // all elementwise arithmetic on packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the kernel was issue bound with its two
    // warps per scheduler (ncu: 58 % issue, 4.6 TB/s)
    float2 ga[NV][4], acc_g[NV][4], acc_b[NV][4], acc_s[HAS_B ? NV : 1][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + (off[i] >> 1));
        const float4 g1 = *reinterpret_cast<const float4*>(p.gamma + (off[i] >> 1) + 4);
        ga[i][0] = make_float2(g0.x, g0.y); ga[i][1] = make_float2(g0.z, g0.w);
        ga[i][2] = make_float2(g1.x, g1.y); ga[i][3] = make_float2(g1.z, g1.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_g[i][j] = acc_b[i][j] = make_float2(0.f, 0.f);
            if (HAS_B) acc_s[HAS_B ? i : 0][j] = make_float2(0.f, 0.f);
        }
    }
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    for (int s = 0; s < stages; ++s)
        fast_issue<NTENS>(ring + (size_t)s * NTENS * row_bytes, &bars[s], src, warp0 + (long long)s * nwarps, p.rows, row_bytes, lane);
    const float inv_c = 1.0f / (float)C;
    const float keep_b = DROP ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    unsigned char* DA = reinterpret_cast<unsigned char*>(p.da);
    unsigned char* DB = reinterpret_cast<unsigned char*>(p.db);
    int it = 0;
    for (long long row = warp0; row < p.rows; row += nwarps, ++it) {
        const int slot = it % stages;
        const float mean = p.mean[row], rstd = p.rstd[row];
        unsigned keep = 0xffffffffu;
        if (HAS_B && DROP) keep = res_keep_bits<NV>(p.seed_b, row, C, off, p.drop_b);
        mbar_wait(&bars[slot], (uint32_t)((it / stages) & 1));
        const unsigned char* sm = ring + (size_t)slot * NTENS * row_bytes;
        uint4 va[NV], vg[NV], vb[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            va[i] = *reinterpret_cast<const uint4*>(sm + off[i]);
            vg[i] = *reinterpret_cast<const uint4*>(sm + row_bytes + off[i]);
            if (HAS_B) vb[i] = *reinterpret_cast<const uint4*>(sm + 2 * row_bytes + off[i]);
        }
        __syncwarp();
        fast_issue<NTENS>(ring + (size_t)slot * NTENS * row_bytes, &bars[slot], src, row + (long long)stages * nwarps, p.rows,
                          row_bytes, lane);
        float2 xh[NV][4], g[NV][4];
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
        const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t wa[4] = {va[i].x, va[i].y, va[i].z, va[i].w}, wg[4] = {vg[i].x, vg[i].y, vg[i].z, vg[i].w};
            const uint32_t wb[4] = {HAS_B ? vb[i].x : 0u, HAS_B ? vb[i].y : 0u, HAS_B ? vb[i].z : 0u, HAS_B ? vb[i].w : 0u};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 x = unpack_bf16x2(wa[j]);
                g[i][j] = unpack_bf16x2(wg[j]);
                if (HAS_B) {
                    float2 t = unpack_bf16x2(wb[j]);
                    if (DROP) {
                        t.x = ((keep >> (8 * i + 2 * j)) & 1u) ? t.x * keep_b : 0.f;
                        t.y = ((keep >> (8 * i + 2 * j + 1)) & 1u) ? t.y * keep_b : 0.f;
                    }
                    x = __fadd2_rn(x, t);
                }
                x = __ffma2_rn(x, rstd2, nmr2);
                xh[i][j] = x;
                acc_g[i][j] = __ffma2_rn(g[i][j], x, acc_g[i][j]);
                acc_b[i][j] = __fadd2_rn(acc_b[i][j], g[i][j]);
                g[i][j] = __fmul2_rn(g[i][j], ga[i][j]);
                s1 = __fadd2_rn(s1, g[i][j]);
                s2 = __ffma2_rn(g[i][j], x, s2);
            }
        }
        const float m1 = warp_sum(s1.x + s1.y) * inv_c;
        const float m2 = warp_sum(s2.x + s2.y) * inv_c;
        const float2 nm2 = make_float2(-m2, -m2), nm1 = make_float2(-m1, -m1);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float2 o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)  // rstd * (g - m1 - xh * m2)
                o[j] = __fmul2_rn(rstd2, __ffma2_rn(xh[i][j], nm2, __fadd2_rn(g[i][j], nm1)));
            if (DA != nullptr) {
                uint4 v;
                v.x = pack_bf16x2(o[0].x, o[0].y); v.y = pack_bf16x2(o[1].x, o[1].y);
                v.z = pack_bf16x2(o[2].x, o[2].y); v.w = pack_bf16x2(o[3].x, o[3].y);
                *reinterpret_cast<uint4*>(DA + row * row_bytes + off[i]) = v;
            }
            if (HAS_B) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (DROP) {
                        o[j].x = ((keep >> (8 * i + 2 * j)) & 1u) ? o[j].x * keep_b : 0.f;
                        o[j].y = ((keep >> (8 * i + 2 * j + 1)) & 1u) ? o[j].y * keep_b : 0.f;
                    }
                    acc_s[HAS_B ? i : 0][j] = __fadd2_rn(acc_s[HAS_B ? i : 0][j], o[j]);
                }
                if (DB != nullptr) {
                    uint4 v;
                    v.x = pack_bf16x2(o[0].x, o[0].y); v.y = pack_bf16x2(o[1].x, o[1].y);
                    v.z = pack_bf16x2(o[2].x, o[2].y); v.w = pack_bf16x2(o[3].x, o[3].y);
                    *reinterpret_cast<uint4*>(DB + row * row_bytes + off[i]) = v;
                }
            }
        }
    }
    // CTA reduction of the column accumulators through shared memory (the rings are idle now), then one global
    // atomic per channel and CTA
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);  // [dgamma | dbeta | dbias_b] x C
    for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = (off[i] >> 1) + j;
            atomicAdd(&red[c], (j & 1) ? acc_g[i][j >> 1].y : acc_g[i][j >> 1].x);
            atomicAdd(&red[C + c], (j & 1) ? acc_b[i][j >> 1].y : acc_b[i][j >> 1].x);
            if (HAS_B) atomicAdd(&red[2 * C + c], (j & 1) ? acc_s[HAS_B ? i : 0][j >> 1].y : acc_s[HAS_B ? i : 0][j >> 1].x);
        }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (p.dgamma != nullptr) atomicAdd(p.dgamma + c, red[c]);
        if (p.dbeta != nullptr) atomicAdd(p.dbeta + c, red[C + c]);
        if (HAS_B && dbias_b != nullptr) atomicAdd(dbias_b + c, red[2 * C + c]);
    }
}

// dispatch of the residual-norm kernels; -1 when the generic path must be used
static int launch_resln_fast(const RowLnParams& p, bool bwd, float* dbias_b, cudaStream_t st) {
    if (p.act != 0 || p.gamma == nullptr || p.post != nullptr || p.drop_out > 0.f) return -1;
    if (p.gw != p.gr || p.C % 256 || p.C < 512 || p.C > 1024) return -1;
    if (p.b == nullptr && p.drop_b > 0.f) return -1;
    if (bwd && (p.b != nullptr) != (p.db != nullptr || dbias_b != nullptr)) return -1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p.a) | reinterpret_cast<uintptr_t>(p.b) | reinterpret_cast<uintptr_t>(p.y) |
                         reinterpret_cast<uintptr_t>(p.dy) | reinterpret_cast<uintptr_t>(p.da) | reinterpret_cast<uintptr_t>(p.db) |
                         reinterpret_cast<uintptr_t>(p.gamma) | reinterpret_cast<uintptr_t>(p.beta);
    if (al & 15) return -1;
    const int nv = p.C / 256;
    const bool has_b = p.b != nullptr, drop = p.drop_b > 0.f;
    const int ntens = bwd ? (has_b ? 3 : 2) : (has_b ? 2 : 1);
    const int row_bytes = p.C * 2;
    const size_t fixed = (bwd ? 0 : (size_t)2 * p.C * sizeof(float)) + (size_t)ROWLN_WARPS * FAST_STAGES_MAX * 8;
    const size_t per_stage = (size_t)ROWLN_WARPS * ntens * row_bytes;
    const size_t budget = bwd ? 200 * 1024 : 110 * 1024;
    int stages = (int)((budget - fixed) / per_stage);
    if (stages > FAST_STAGES_MAX) stages = FAST_STAGES_MAX;
    if (stages < 2) return -1;
    size_t smem = per_stage * stages + fixed;
    if (bwd && smem < (size_t)3 * p.C * sizeof(float)) smem = (size_t)3 * p.C * sizeof(float);
    const long long blocks_needed = ceil_div64(p.rows, ROWLN_WARPS);
    const long long cap = (long long)a2v_num_sms() * (bwd ? 1 : 2);
    int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    if (grid < 1) grid = 1;
#define A2V_RES(KF, KB)                                                                                          \
    do {                                                                                                         \
        if (a2v_ensure_dynamic_smem(bwd ? reinterpret_cast<const void*>(KB) : reinterpret_cast<const void*>(KF), smem) !=  \
            A2V_OK)                                                                                              \
            return A2V_ERR_CUDA;                                                                                 \
        if (bwd) KB<<<grid, ROWLN_WARPS * 32, smem, st>>>(p, stages, dbias_b);                                    \
        else KF<<<grid, ROWLN_WARPS * 32, smem, st>>>(p, stages);                                                 \
    } while (0)
#define A2V_RES_NV(N)                                                                                   \
    do {                                                                                                \
        if (has_b && drop) A2V_RES((resln_fwd_kernel<N, true, true>), (resln_bwd_kernel<N, true, true>)); \
        else if (has_b) A2V_RES((resln_fwd_kernel<N, true, false>), (resln_bwd_kernel<N, true, false>));  \
        else A2V_RES((resln_fwd_kernel<N, false, false>), (resln_bwd_kernel<N, false, false>));           \
    } while (0)
    switch (nv) {
        case 2: A2V_RES_NV(2); break;
        case 3: A2V_RES_NV(3); break;
        default: A2V_RES_NV(4); break;
    }
#undef A2V_RES
#undef A2V_RES_NV
    return a2v_check_launch(bwd ? "rowln_bwd(res)" : "rowln_fwd(res)");
}

// dispatch test + launch of the specialised kernels; returns -1 when the generic path must be used
static int launch_rowln_fast(const RowLnParams& p, bool bwd, cudaStream_t st) {
    if (p.act != 1 || p.gamma != nullptr || p.beta != nullptr || p.b != nullptr) return -1;
    if (p.drop_out > 0.f) return -1;
    if (bwd && (p.da == nullptr || p.db != nullptr)) return -1;
    if (p.gw % 8 || p.gr % 8 || p.C % p.gw) return -1;
    ChunkMap cm;
    cm.cs = p.gw / 8; cm.cr = p.gr / 8; cm.ngroups = p.C / p.gw;
    const int real_chunks = cm.cr * cm.ngroups;
    if (real_chunks % 32 || real_chunks / 32 < 1 || real_chunks / 32 > 4) return -1;
    const int nv = real_chunks / 32;
    const int ntens = bwd ? 2 : (p.post != nullptr ? 2 : 1);
    const int row_bytes = p.C * 2;
    const size_t per_stage = (size_t)ROWLN_WARPS * ntens * row_bytes;
    int stages = (int)((110 * 1024 - ROWLN_WARPS * FAST_STAGES_MAX * 8) / per_stage);
    if (stages > FAST_STAGES_MAX) stages = FAST_STAGES_MAX;
    if (stages < 2) return -1;
    const size_t smem = per_stage * stages + (size_t)ROWLN_WARPS * FAST_STAGES_MAX * 8;
    long long blocks_needed = ceil_div64(p.rows, ROWLN_WARPS);
    long long cap = (long long)a2v_num_sms() * 2;
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
#define A2V_FAST(KERNEL)                                                                                         \
    do {                                                                                                         \
        auto k = KERNEL;                                                                                         \
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(k), smem) != A2V_OK) return A2V_ERR_CUDA;      \
        k<<<grid, ROWLN_WARPS * 32, smem, st>>>(p, stages, cm);                                                   \
    } while (0)
#define A2V_FAST_NV(N)                                                          \
    do {                                                                        \
        if (bwd) A2V_FAST((rowln_gelu_bwd_kernel<N>));                          \
        else if (p.post != nullptr) A2V_FAST((rowln_gelu_fwd_kernel<N, true>)); \
        else A2V_FAST((rowln_gelu_fwd_kernel<N, false>));                       \
    } while (0)
    switch (nv) {
        case 1: A2V_FAST_NV(1); break;
        case 2: A2V_FAST_NV(2); break;
        case 3: A2V_FAST_NV(3); break;
        default: A2V_FAST_NV(4); break;
    }
#undef A2V_FAST
#undef A2V_FAST_NV
    return a2v_check_launch(bwd ? "rowln_bwd(fast)" : "rowln_fwd(fast)");
}

template <typename T>
static int launch_rowln(const RowLnParams& p, bool bwd, cudaStream_t st) {
    const int nch = ceil_div(p.C, 128);
    const bool full = (p.gr == p.gw) && (p.C == 128 || p.C == 256 || p.C == 512 || p.C == 1024);
    const int threads = ROWLN_WARPS * 32;
    const int row_bytes = p.C * (int)sizeof(T);
    const int ntens = bwd ? (p.b != nullptr ? 3 : 2) : (1 + (p.b != nullptr ? 1 : 0) + (p.post != nullptr ? 1 : 0));
    const size_t fixed = (size_t)(bwd ? 8 : 4) * p.C * sizeof(float) + (size_t)ROWLN_WARPS * 8 * 8;
    const size_t per_stage = (size_t)ROWLN_WARPS * ntens * row_bytes;
    // forward kernels fit two CTAs per SM (about 104 registers): keep each under ~110 KB; the
    // backward kernel is register-limited to one CTA per SM and may take up to ~200 KB.
    const bool one_cta = bwd && p.gamma != nullptr;  // affine backward: ~220 registers (parameter-gradient accumulators)
    const size_t budget = one_cta ? 200 * 1024 : 110 * 1024;
    int stages = (int)((budget > fixed ? budget - fixed : 0) / per_stage);
    if (stages < 2) stages = 2;
    if (stages > 6) stages = 6;
    const size_t smem = fixed + per_stage * stages;
    A2V_REQUIRE(smem <= 227 * 1024, "rowln: row of %d channels does not fit the shared-memory pipeline", p.C);
    const int ctas_per_sm = one_cta ? 1 : (smem <= 113 * 1024 ? 2 : 1);
    long long blocks_needed = ceil_div64(p.rows, ROWLN_WARPS);
    long long cap = (long long)a2v_num_sms() * ctas_per_sm;
    int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    if (grid < 1) grid = 1;
#define A2V_ROWLN_K(N, F, A, AF)                                                                           \
    do {                                                                                                  \
        auto kf = rowln_fwd_kernel<T, N, F, A, AF>;                                                       \
        auto kb = rowln_bwd_kernel<T, N, F, A, AF>;                                                       \
        if (a2v_ensure_dynamic_smem(bwd ? reinterpret_cast<const void*>(kb) : reinterpret_cast<const void*>(kf), smem) != \
            A2V_OK)                                                                                       \
            return A2V_ERR_CUDA;                                                                          \
        if (bwd)                                                                                          \
            kb<<<grid, threads, smem, st>>>(p, stages);                                                   \
        else                                                                                              \
            kf<<<grid, threads, smem, st>>>(p, stages);                                                   \
    } while (0)
#define A2V_ROWLN_F(N, F)                                                       \
    do {                                                                        \
        if (p.act == 2) A2V_ROWLN_K(N, false, 2, true);                         \
        else if (p.act == 1 && affine_) A2V_ROWLN_K(N, F, 1, true);             \
        else if (p.act == 1) A2V_ROWLN_K(N, F, 1, false);                       \
        else if (affine_) A2V_ROWLN_K(N, F, 0, true);                           \
        else A2V_ROWLN_K(N, F, 0, false);                                       \
    } while (0)
#define A2V_ROWLN(N)                                  \
    do {                                              \
        if (full) A2V_ROWLN_F(N, true);               \
        else A2V_ROWLN_F(N, false);                   \
    } while (0)
    const bool affine_ = p.gamma != nullptr;
    A2V_REQUIRE(p.act != 2 || (affine_ && !full), "rowln: PSwish is built for the affine, padded 127-of-128 case");
    A2V_REQUIRE(affine_ || p.beta == nullptr, "rowln: beta without gamma");
    if (nch <= 1) A2V_ROWLN(1);
    else if (nch <= 2) A2V_ROWLN(2);
    else if (nch <= 4) A2V_ROWLN(4);
    else A2V_ROWLN(8);
#undef A2V_ROWLN_K
#undef A2V_ROWLN_F
#undef A2V_ROWLN
    return a2v_check_launch(bwd ? "rowln_bwd" : "rowln_fwd");
}

static int validate_rowln(const a2v_rowln_desc* d) {
    A2V_REQUIRE(d != nullptr, "rowln: NULL descriptor");
    A2V_REQUIRE(d->rows >= 0 && d->channels > 0 && d->channels % 8 == 0 && d->channels <= 1024,
                "rowln: channels must be a multiple of 8 in (0, 1024], got %d", d->channels);
    A2V_REQUIRE(d->group_width > 0 && d->group_real > 0 && d->group_real <= d->group_width &&
                    d->channels % d->group_width == 0,
                "rowln: bad group padding (%d real of %d, C=%d)", d->group_real, d->group_width, d->channels);
    A2V_REQUIRE(d->dtype == A2V_F32 || d->dtype == A2V_BF16, "rowln: bad dtype");
    A2V_REQUIRE(d->a != nullptr, "rowln: a is NULL");
    A2V_REQUIRE(d->act >= 0 && d->act <= 2, "rowln: act must be 0, 1 or 2");
    A2V_REQUIRE(d->act != 2 || (d->act_alpha != nullptr && d->act_beta != nullptr), "rowln: PSwish needs alpha/beta");
    A2V_REQUIRE(d->drop_b >= 0.f && d->drop_b < 1.f && d->drop_out >= 0.f && d->drop_out < 1.f,
                "rowln: dropout probabilities must be in [0, 1)");
    return A2V_OK;
}

static RowLnParams to_params(const a2v_rowln_desc* d) {
    RowLnParams p;
    p.a = d->a; p.b = d->b; p.gamma = d->gamma; p.beta = d->beta;
    p.act_alpha = d->act_alpha; p.act_beta = d->act_beta; p.post = d->post;
    p.y = d->y; p.mean = d->mean; p.rstd = d->rstd;
    p.rows = d->rows; p.C = d->channels; p.gw = d->group_width; p.gr = d->group_real;
    p.eps = d->eps; p.act = d->act;
    p.drop_b = d->drop_b; p.seed_b = d->seed_b; p.drop_out = d->drop_out; p.seed_out = d->seed_out;
    p.dy = d->dy; p.da = d->da; p.db = d->db;
    p.dgamma = d->dgamma; p.dbeta = d->dbeta; p.dact_alpha = d->dact_alpha; p.dact_beta = d->dact_beta;
    return p;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_rowln_fwd(const a2v_rowln_desc* d, a2v_stream_t stream) {
    int rc = validate_rowln(d);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->y != nullptr, "rowln_fwd: y is NULL");
    if (d->rows == 0) return A2V_OK;
    RowLnParams p = to_params(d);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d->dtype == A2V_BF16) {
        rc = launch_rowln_fast(p, false, st);
        if (rc >= 0) return rc;
        rc = launch_resln_fast(p, false, nullptr, st);
        if (rc >= 0) return rc;
        rc = launch_rowln128(p, false, st);
        if (rc >= 0) return rc;
    }
    return d->dtype == A2V_F32 ? launch_rowln<float>(p, false, st) : launch_rowln<bf16>(p, false, st);
}

extern "C" int a2v_rowln_bwd(const a2v_rowln_desc* d, a2v_stream_t stream) {
    int rc = validate_rowln(d);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->dy != nullptr && d->mean != nullptr && d->rstd != nullptr,
                "rowln_bwd: dy / saved mean / saved rstd are required");
    A2V_REQUIRE(d->db == nullptr || d->b != nullptr, "rowln_bwd: db requested without b");
    A2V_REQUIRE(d->dbias_b == nullptr || d->db != nullptr, "rowln_bwd: dbias_b (column sums of db) requested without db");
    if (d->rows == 0) return A2V_OK;
    RowLnParams p = to_params(d);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d->dtype == A2V_BF16) {
        rc = launch_rowln_fast(p, true, st);
        if (rc >= 0) return rc;
        rc = launch_resln_fast(p, true, d->dbias_b, st);
        if (rc >= 0) return rc;
        rc = launch_rowln128(p, true, st);
        if (rc >= 0) return rc;
    }
    rc = d->dtype == A2V_F32 ? launch_rowln<float>(p, true, st) : launch_rowln<bf16>(p, true, st);
    if (rc != A2V_OK || d->dbias_b == nullptr) return rc;
    return a2v_colsum(d->dtype, d->db, d->dbias_b, d->rows, d->channels, stream);  // generic path: separate reduction
}
