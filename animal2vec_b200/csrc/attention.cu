// Multi-head self-attention with on-the-fly ALiBi (head dim 64).
//
// Reference behaviour replaced (file:line under /root/reference):
//   nn/modalities/modules.py:368-410 AltAttention.forward: q*scale @ k^T, + alibi bias (fp32),
//     softmax fp32, dropout, @ v.
//   nn/modalities/base.py:553-698 get_alibi / get_alibi_bias / masked_alibi and base.py:293-314
//     (per-head learned scale, clamp_min 0, clone repeat, double gather by ids_keep): the
//     (B*M, H, T, T) fp32 bias tensor is never built; bias[h,i,j] = -slope_h*max(scale_h,0)*|pos_i-pos_j|
//     is evaluated from the token positions inside the kernels.
//
// Kernels:
//   attn_fwd_tcgen05_kernel : bf16, flash-style. S = Q K^T and O_j = P V_j run on tcgen05 with
//                             TMEM accumulators, Q/K/V tiles arrive by TMA, one thread per query row
//                             does the online softmax straight out of TMEM (no shuffles).
//   attn_bwd_tcgen05_kernel : bf16, student shapes (L <= 160 kept tokens): persistent per head, whole head resident
//                             in shared memory, all five products on tcgen05 with TMEM accumulators.
//   attn_qk_bound_kernel    : max |q|, max |k| per head (the data-dependent part of the ALiBi key-tile window).
//   attn_*_ref_kernel       : fp32 CUDA-core kernels for the fp32 validation mode.
#include <stdlib.h>
#include "attention_common.cuh"

namespace a2v {

// ------------------------------------------------------------------------------------------
// forward, tcgen05
// ------------------------------------------------------------------------------------------
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_K = 16384;             // 2 buffers
constexpr int ATT_SMEM_V = 16384 * 3;         // 2 buffers
constexpr int ATT_SMEM_POS = 16384 * 5;       // 128 floats (key positions of the current tile)
// ALiBi inside the score product (contiguous sequences): three 128 x 16 bf16 operands without swizzle (8 x 8 core
// matrices of 128 B; K-adjacent ones 128 B apart, 8-row groups 256 B apart) -- the key-column index, and +/- the slope
// split into three bf16 pieces
constexpr int ATT_SMEM_EXT = ATT_SMEM_POS + 512;
constexpr int ATT_EXT_BYTES = 4096;
constexpr int ATT_SMEM_BAR = ATT_SMEM_EXT + 3 * ATT_EXT_BYTES;
// 92.6 KB (the probabilities live in tensor memory, not in shared memory): two CTAs per SM. The dynamic window
// of a kernel without static shared memory starts 1024-aligned (checked at run time, trap otherwise).
constexpr int ATT_SMEM_TOTAL = ATT_SMEM_BAR + 128;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units: the running maximum may lag by up to 2^8
// ALiBi locality: a key tile is skipped when EVERY probability in it is provably below 2^-50 of its row's largest
// term (its whole tile adds < 2^-39 relative to the fp32 row sum, 2^15 below fp32 resolution -- the reference's own
// fp32 softmax, nn/modalities/modules.py:396-399, rounds such terms away as well).
constexpr float ATT_SKIP_LOG2 = 50.0f;


// Flash-style forward. One thread per query row; per 128-key tile:
//   S = Q K^T (tcgen05, TMEM)  ->  ONE TMEM read into registers, ALiBi + running max in log2 units
//   ->  P = exp2(S - m) as packed bf16 into TMEM  ->  O += P V accumulated IN TMEM (tcgen05, A operand from TMEM).
// O is rescaled in TMEM only when a row's maximum grows by more than 2^8 (lazy rescaling), the S MMA of
// the next tile is issued before the softmax of this one finishes, K and V tiles have separate
// barriers so the next K can land while V is still being consumed.
// Contiguous sequences (ncu, profiles/r2_ncu_attn_teacher.md: the kernel is co-limited by issue slots and the MUFU
// queue, ~50 % each, 895 warp instructions per 128 x 128 tile):
//  * the key tiles are visited diagonal first, then outwards: the running maximum is (nearly) final after the first
//    tile, so the O accumulator is not rescaled on the way towards the diagonal;
//  * off the diagonal |i - j| is linear in the key column, and that term rides in the score product as a fifth K step
//    (A = the slope in three bf16 pieces, B = the column index; exact products, fp32 accumulation) -- no per-element
//    FFMA for the bias;
//  * the exponent arguments and the row sums use packed fp32 pairs (FFMA2 / FADD2).
#ifdef A2V_ATTN_TRACE
__device__ long long g_ft_trace[64];
#define FT_TR(k) do { if (blockIdx.x == 5 && blockIdx.y == 3 && blockIdx.z == 1 && j == 6 && tid == 0) g_ft_trace[k] = clock64(); } while (0)
#else
#define FT_TR(k) do { } while (0)
#endif

template <bool HAS_POS, bool DROP, bool TRIM>
__global__ void __launch_bounds__(128, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0u) __trap();  // the 128B-swizzled tiles need 1024-byte alignment
    // second launch after attention_stream.cu: only the heads that kernel declined
    if (!HAS_POS && p.head_filter == 1 && attn_stream_head_ok(p, blockIdx.z, blockIdx.y)) return;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
    uint64_t* bar_q = bars;
    uint64_t* bar_k = bars + 1;  // [2]
    uint64_t* bar_v = bars + 3;  // [2]
    uint64_t* bar_s = bars + 5;
    uint64_t* bar_o = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    float* spos = reinterpret_cast<float*>(smem + ATT_SMEM_POS);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int h = blockIdx.y, b = blockIdx.z;
    const int L = p.L, D = p.D;
    const int n_kv = (L + 127) >> 7;

    if (tid == 0) {
        tma_prefetch_desc(&tm);
        for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    const float kappa = head_coef(p, h) / p.sm_scale;  // ALiBi slope in units of the raw q.k product
    const bool use_ext = !HAS_POS && kappa != 0.f;
    if (use_ext) {
        const __nv_bfloat16 k1 = __float2bfloat16_rn(kappa);
        const __nv_bfloat16 k2 = __float2bfloat16_rn(kappa - __bfloat162float(k1));
        const __nv_bfloat16 k3 = __float2bfloat16_rn(kappa - __bfloat162float(k1) - __bfloat162float(k2));
        const uint32_t b1 = __bfloat16_as_ushort(k1), b2 = __bfloat16_as_ushort(k2), b3 = __bfloat16_as_ushort(k3);
        const uint32_t bc = __bfloat16_as_ushort(__float2bfloat16_rn((float)tid));  // 0..127: exact
        uint8_t* ext = smem + ATT_SMEM_EXT + (tid >> 3) * 256 + (tid & 7) * 16;      // row tid, K elements 0..7
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(ext) = make_uint4(bc | (bc << 16), bc, 0u, 0u);                                  // key column
        *reinterpret_cast<uint4*>(ext + ATT_EXT_BYTES) = make_uint4(b1 | (b2 << 16), b3, 0u, 0u);                  // + slope
        *reinterpret_cast<uint4*>(ext + 2 * ATT_EXT_BYTES) = make_uint4((b1 | (b2 << 16)) ^ 0x80008000u, b3 ^ 0x8000u, 0u, 0u);  // - slope
        *reinterpret_cast<uint4*>(ext + 128) = zero;                                                               // K elements 8..15
        *reinterpret_cast<uint4*>(ext + ATT_EXT_BYTES + 128) = zero;
        *reinterpret_cast<uint4*>(ext + 2 * ATT_EXT_BYTES + 128) = zero;
        fence_proxy_async();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;        // 128 columns
    const uint32_t tmem_o = tmem_base + 128;  // 64 columns
    const uint32_t tmem_p = tmem_base + 192;  // 64 columns: P as packed bf16 pairs, the TMEM A operand of P.V
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);
    const uint32_t qa = smem_u32(smem + ATT_SMEM_Q);

    // One query tile per CTA (grid.x = number of query tiles); the head-filtered second launch behind attention_stream.cu
    // uses grid.x = 1 and walks the query tiles of its (rarely taken) heads here, so a declined launch costs
    // batch x heads empty CTAs instead of batch x heads x tiles.
    for (int qt = blockIdx.x; qt < n_kv; qt += gridDim.x) {
    const int q0 = qt * 128;
    // Key-tile range [j_begin, j_end) of this query tile. Contiguous sequences with an ALiBi slope: with
    // B = max|q| max|k| of the head, every score obeys |q.k| * scale2 <= B * scale2, the row's own key (distance 0)
    // floors the running maximum at -B * scale2, so a key at distance d has exponent <= 2 B scale2 - coef2 d (log2).
    // Tiles whose nearest key is at least w = (2 B scale2 + ATT_SKIP_LOG2) / coef2 frames away contribute nothing.
    int j_begin = 0, j_end = n_kv;
    if (!HAS_POS && p.qk_bound != nullptr) {
        const float c2 = head_coef(p, h) * LOG2E;
        if (c2 > 0.f) {
            const float2 b2 = reinterpret_cast<const float2*>(p.qk_bound)[b * p.H + h];  // max|q|^2, max|k|^2
            const float qk = sqrtf(b2.x) * sqrtf(b2.y) * 1.002f;
            const float w = (2.f * qk * (p.sm_scale * LOG2E) + ATT_SKIP_LOG2) / c2;
            if (w < 1.0e6f) {
                const int wi = (int)w + 1;
                const int lo = q0 - 127 - wi;  // tile j is kept iff 128 j > lo and 128 j < q0 + 127 + wi
                j_begin = lo < 0 ? 0 : lo / 128 + 1;
                const int hi = (q0 + 126 + wi) / 128 + 1;
                j_end = hi < n_kv ? hi : n_kv;
            }
        }
    }
    const int n_it = j_end - j_begin;

    // key tile of iteration jj: ascending for token positions; contiguous sequences start on the diagonal tile
    // (= this CTA's query tile) and walk outwards, left side first
    const int diag = qt;
    auto tile_of = [&](int jj) -> int {
        if (HAS_POS) return j_begin + jj;
        const int nl = diag - j_begin;
        return jj == 0 ? diag : (jj <= nl ? diag - jj : j_begin + jj);
    };
    // tiles whose ALiBi term is linear in the key column: everything but the diagonal and the ragged last tile
    auto is_linear = [&](int j) -> bool { return !HAS_POS && j != n_kv - 1 && j != diag; };
    auto issue_s = [&](int buf, int j) {
        const uint32_t ka = smem_u32(smem + ATT_SMEM_K + buf * 16384);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tmem_s, umma_smem_desc(qa + k * 32, 0, 1024), umma_smem_desc(ka + k * 32, 0, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
        if (use_ext && is_linear(j)) {  // + sg * slope * column: keys before the query rows count up, keys after count down
            const uint32_t ea = smem_u32(smem + ATT_SMEM_EXT);
            umma_bf16(tmem_s, umma_smem_desc_nosw(ea + (j < diag ? 1 : 2) * ATT_EXT_BYTES, 128, 256),
                      umma_smem_desc_nosw(ea, 128, 256), idesc_s, 1u);
        }
        umma_commit(bar_s);
    };

    if (warp == 0 && elect_one()) {
        mbar_expect_tx(bar_q, 16384);
        tma_load_3d(smem + ATT_SMEM_Q, &tm, bar_q, h * HD, q0, b);
        mbar_expect_tx(&bar_k[0], 16384);
        tma_load_3d(smem + ATT_SMEM_K, &tm, &bar_k[0], D + h * HD, tile_of(0) * 128, b);
        mbar_expect_tx(&bar_v[0], 16384);
        tma_load_3d(smem + ATT_SMEM_V, &tm, &bar_v[0], 2 * D + h * HD, tile_of(0) * 128, b);
        if (n_it > 1) {
            mbar_expect_tx(&bar_k[1], 16384);
            tma_load_3d(smem + ATT_SMEM_K + 16384, &tm, &bar_k[1], D + h * HD, tile_of(1) * 128, b);
        }
        mbar_wait(bar_q, 0);
        mbar_wait(&bar_k[0], 0);
        tc_fence_after();
        issue_s(0, tile_of(0));
    }

    const int qi = q0 + tid;  // this thread's query row
    const bool q_ok = qi < L;
    const int pos_i = q_ok ? (HAS_POS ? p.pos[(long long)b * L + qi] : qi) : 0;
    const float coef2 = head_coef(p, h) * LOG2E;
    const float scale2 = p.sm_scale * LOG2E;
    const float inv_keep = DROP ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const long long bh = (long long)b * p.H + h;
    const uint32_t row_key = DROP ? attn_row_key(p.seed, bh, L, qi) : 0u;
    const uint32_t drop_thr = attn_drop_threshold(p.drop_p);

    float m_run = -INFINITY, l_run = 0.f;
    // warps whose 32 query rows all lie beyond L (second q tile of a short sequence) only keep the barriers moving
    const bool warp_active = !TRIM || (q0 + warp * 32) < L;  // TRIM: short sequences (student), skip dead work

    for (int jj = 0; jj < n_it; ++jj) {  // jj: iteration (buffers, barrier phases); j: key tile (positions)
        const int j = tile_of(jj);
        const int buf = jj & 1;
        const int k0 = j * 128;
        const bool last = (j == n_kv - 1);
        if (HAS_POS) {
            const int kj = k0 + tid;
            // single buffer: the __syncthreads before the P.V issue of the previous tile orders its readers
            spos[tid] = kj < L ? (float)p.pos[(long long)b * L + kj] : -268435456.0f;
            __syncthreads();
        }
        FT_TR(0);
        mbar_wait(bar_s, jj & 1);
        tc_fence_after();
        FT_TR(1);
        if (warp == 0 && elect_one() && jj + 2 < n_it) {  // K buffer `buf` is free: S(j) has been computed
            mbar_expect_tx(&bar_k[buf], 16384);
            tma_load_3d(smem + ATT_SMEM_K + buf * 16384, &tm, &bar_k[buf], D + h * HD, tile_of(jj + 2) * 128, b);
        }

        // scores -> registers, row maximum. t[] holds u with  score(log2 units) = u * e_mul + c_row:
        //  * generic tiles (token positions, the diagonal tile, the ragged last tile): u = score, e_mul = 1, c_row = 0;
        //  * every other tile of a contiguous sequence lies entirely before or after this CTA's query rows, so
        //    |i - j| is linear in the key column and already inside the product: score = scale2 * raw' + c_row with a
        //    per-row constant c_row that folds into the exp2 offset -- half an FMNMX3 per element in this pass.
        float t[128];
        float m_tile = -INFINITY;
        float e_mul = 1.0f, c_row = 0.f;
        const float dist0 = (float)(pos_i - k0);
        const float fpos_i = (float)pos_i;
        const int nvalid = L - k0;  // keys of this tile that exist (>= 128 except for the last tile)
        if (is_linear(j)) {
            // the product already holds raw + sg * slope * column (fifth K step of issue_s)
            const float sg = j < diag ? 1.0f : -1.0f;
            float m_u = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_s + lane_off + c * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float u = __uint_as_float(raw[i]);
                    t[c * 32 + i] = u;
                    m_u = fmaxf(m_u, u);
                }
            }
            e_mul = scale2;
            c_row = -sg * coef2 * dist0;
            m_tile = fmaf(m_u, scale2, c_row);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (TRIM && (!warp_active || c * 32 >= nvalid)) {  // nothing real in this 32-key chunk
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[c * 32 + i] = -INFINITY;
                    continue;
                }
                uint32_t raw[32];
                tmem_ld_32x32(tmem_s + lane_off + c * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = c * 32 + i;
                    float dist;
                    if (HAS_POS) dist = fpos_i - spos[col];
                    else dist = dist0 - (float)col;
                    float v = fmaf(__uint_as_float(raw[i]), scale2, -coef2 * fabsf(dist));
                    if (last && col >= nvalid) v = -INFINITY;
                    t[col] = v;
                    m_tile = fmaxf(m_tile, v);
                }
            }
        }
        FT_TR(2);
        // every thread holds its scores: the S accumulator is free, so S of the NEXT tile runs on the tensor
        // pipe while this tile's exponentials are computed (it used to be issued after P.V, leaving the CTA
        // waiting a full MMA round trip at the top of every iteration)
        if (jj + 1 < n_it) {
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                mbar_wait(&bar_k[buf ^ 1], ((jj + 1) >> 1) & 1);
                tc_fence_after();
                FT_TR(8);
                issue_s(buf ^ 1, tile_of(jj + 1));
                FT_TR(9);
#ifdef A2V_ATTN_TRACE
                if (blockIdx.x == 5 && blockIdx.y == 3 && blockIdx.z == 1 && j == 6) {
                    while (!mbar_try_wait(bar_s, (jj + 1) & 1)) {}
                    g_ft_trace[10] = clock64();
                }
#endif
            }
        }
        // previous P.V done: P smem, V[buf^1] and the O accumulator are ours again
        if (jj > 0) {
            mbar_wait(bar_o, (jj - 1) & 1);
            tc_fence_after();
        }
        FT_TR(3);
        if (warp == 0 && elect_one() && jj + 1 < n_it) {
            mbar_expect_tx(&bar_v[buf ^ 1], 16384);
            tma_load_3d(smem + ATT_SMEM_V + (buf ^ 1) * 16384, &tm, &bar_v[buf ^ 1], 2 * D + h * HD, tile_of(jj + 1) * 128, b);
        }
        // lazy rescaling of O (in TMEM) and of the running sum
        const bool grow = m_tile > m_run + ATT_RESCALE_THRESHOLD;  // also true on the first tile (m_run = -inf)
        if (jj > 0 && warp_active && __any_sync(0xffffffffu, grow)) {
            const float alpha = grow ? ex2_approx(m_run - m_tile) : 1.0f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_o + lane_off + c * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
                tmem_st_32x32(tmem_o + lane_off + c * 32, raw);
            }
            tmem_st_wait();
            l_run *= alpha;
        }
        if (grow) m_run = m_tile;

        // probabilities -> TENSOR memory (packed bf16 pairs, 16 columns per 32 keys): the A operand of P.V is read
        // from TMEM, so P costs no shared-memory bandwidth (S and P.V operands out of smem were the bottleneck:
        // 144 KB per 128x128 tile against 128 B/clk)
        float2 l2 = make_float2(0.f, 0.f);
        const float e_off = c_row - m_run;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (TRIM && !warp_active) continue;  // rows beyond L: their P rows only feed O rows that are never stored
            uint32_t pk[16];
            if (TRIM && c * 32 >= nvalid) {      // keys beyond L: exact zeros (they sit inside the K extent of P.V)
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {  // packed fp32 pairs: one FFMA2 / FADD2 per two keys
                        const float2 a = __ffma2_rn(make_float2(t[c * 32 + u * 8 + i], t[c * 32 + u * 8 + i + 1]),
                                                    make_float2(e_mul, e_mul), make_float2(e_off, e_off));
                        e[i] = ex2_approx(a.x);
                        e[i + 1] = ex2_approx(a.y);
                        l2 = __fadd2_rn(l2, make_float2(e[i], e[i + 1]));
                    }
                    if (DROP) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const uint2 bits = attn_bits4(row_key, (k0 + c * 32 + u * 8) / 4 + g);
                            e[4 * g + 0] = (bits.x & 0xffffu) >= drop_thr ? e[4 * g + 0] * inv_keep : 0.f;
                            e[4 * g + 1] = (bits.x >> 16) >= drop_thr ? e[4 * g + 1] * inv_keep : 0.f;
                            e[4 * g + 2] = (bits.y & 0xffffu) >= drop_thr ? e[4 * g + 2] * inv_keep : 0.f;
                            e[4 * g + 3] = (bits.y >> 16) >= drop_thr ? e[4 * g + 3] * inv_keep : 0.f;
                        }
                    }
                    pk[u * 4 + 0] = pack_bf16x2(e[0], e[1]);
                    pk[u * 4 + 1] = pack_bf16x2(e[2], e[3]);
                    pk[u * 4 + 2] = pack_bf16x2(e[4], e[5]);
                    pk[u * 4 + 3] = pack_bf16x2(e[6], e[7]);
                }
            }
            tmem_st_32x16(tmem_p + lane_off + c * 16, pk);
        }
        const float l_tile = l2.x + l2.y;
        tmem_st_wait();
        l_run += l_tile;
        FT_TR(4);
        tc_fence_before();
        __syncthreads();
        FT_TR(5);
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            FT_TR(11);
            mbar_wait(&bar_v[buf], (jj >> 1) & 1);
            FT_TR(12);
            const uint32_t va = smem_u32(smem + ATT_SMEM_V + buf * 16384);
#pragma unroll
            for (int k = 0; k < 8; ++k)  // 16 keys = 8 packed columns of P per step
                umma_bf16_ts(tmem_o, tmem_p + k * 8, umma_smem_desc(va + k * 2048, 8192, 1024), idesc_o,
                             (jj > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_o);
            FT_TR(13);
        }
        FT_TR(6);
    }

    mbar_wait(bar_o, (n_it - 1) & 1);
    tc_fence_after();
    {
        const float inv_l = 1.0f / l_run;
        bf16* orow = reinterpret_cast<bf16*>(p.out) + ((long long)b * L + qi) * D + h * HD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem_o + lane_off + c * 32, raw);
            tmem_ld_wait();
            if (q_ok) {
#pragma unroll
                for (int d = 0; d < 32; d += 4) {
                    float v[4] = {__uint_as_float(raw[d]) * inv_l, __uint_as_float(raw[d + 1]) * inv_l,
                                  __uint_as_float(raw[d + 2]) * inv_l, __uint_as_float(raw[d + 3]) * inv_l};
                    store4(orow + c * 32 + d, v);
                }
            }
        }
        if (q_ok && p.lse != nullptr) p.lse[bh * L + qi] = (m_run + log2f(l_run)) * LN2;
    }
    if (qt + (int)gridDim.x < n_kv) {  // next query tile of this head: every barrier is quiescent, start their phases over
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < 7; ++i) {
                mbar_inval(&bars[i]);
                mbar_init(&bars[i], 1);
            }
            mbar_fence_init();
            fence_proxy_async();
        }
        __syncthreads();
        tc_fence_after();
    }
    }  // query tiles
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// bound2[(b * H + h) * 2 + {0, 1}] = max_i |q_i|^2, max_j |k_j|^2: the data-dependent part of the ALiBi locality window
// of attn_fwd_tcgen05_kernel. Eight lanes per row (16 bytes each), grid (H, batch, row chunks); the chunks combine with
// an integer atomicMax (non-negative floats order like their bit patterns); the caller zero-fills the buffer.
constexpr int QKB_CHUNK = 256;  // rows per block
__global__ void __launch_bounds__(256) attn_qk_bound_kernel(const bf16* __restrict__ qkv, float* __restrict__ bound2,
                                                            int L, int H) {
    __shared__ float red[2][8];
    const int h = blockIdx.x, b = blockIdx.y, D = H * HD;
    const int sub = threadIdx.x & 7;
    const int r0 = blockIdx.z * QKB_CHUNK;
    float mq = 0.f, mk = 0.f;
#pragma unroll
    for (int it = 0; it < QKB_CHUNK / 32; ++it) {  // uniform trip count: the shuffles below need every lane of the warp
        const int i = r0 + it * 32 + (threadIdx.x >> 3);
        uint4 vq = make_uint4(0u, 0u, 0u, 0u), vk = make_uint4(0u, 0u, 0u, 0u);
        if (i < L) {
            const bf16* row = qkv + ((long long)b * L + i) * 3 * D + h * HD + sub * 8;
            vq = *reinterpret_cast<const uint4*>(row);
            vk = *reinterpret_cast<const uint4*>(row + D);
        }
        float sq = 0.f, sk = 0.f;
        const uint32_t wq[4] = {vq.x, vq.y, vq.z, vq.w}, wk[4] = {vk.x, vk.y, vk.z, vk.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 a = unpack_bf16x2(wq[c]), k2 = unpack_bf16x2(wk[c]);
            sq = fmaf(a.x, a.x, fmaf(a.y, a.y, sq));
            sk = fmaf(k2.x, k2.x, fmaf(k2.y, k2.y, sk));
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            sk += __shfl_xor_sync(0xffffffffu, sk, o);
        }
        mq = fmaxf(mq, sq);
        mk = fmaxf(mk, sk);
    }
    mq = warp_max(mq);
    mk = warp_max(mk);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = mq;
        red[1][threadIdx.x >> 5] = mk;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            mq = fmaxf(mq, red[0][w]);
            mk = fmaxf(mk, red[1][w]);
        }
        atomicMax(reinterpret_cast<int*>(bound2) + (b * H + h) * 2, __float_as_int(mq));
        atomicMax(reinterpret_cast<int*>(bound2) + (b * H + h) * 2 + 1, __float_as_int(mk));
    }
}

constexpr int BWD_LMAX = 160;  // tokens per sequence the shared-memory-resident backward supports

// ------------------------------------------------------------------------------------------
// backward, tcgen05 (student shapes: L <= 160 kept tokens), persistent over (batch, head)
//
// Per head all five products run on tcgen05 with TMEM accumulators; Q, K, V, dO are TMA-loaded ONCE
// as [row][64] 128B-swizzled tiles and serve as K-major or MN-major operands as each product needs:
//   S  = Q K^T          (A = Q K-major,  B = K K-major,  N = 160)   -> P = exp(S*scale + alibi - lse)
//   dP = dO V^T         (A = dO K-major, B = V K-major)              -> dS = P * (dP*keep - delta)
//   dV += Pd^T dO       (A = Pd MN-major, B = dO MN-major)     Pd = P with the dropout mask applied
//   dK += dS^T Q        (A = dS MN-major, B = Q MN-major)
//   dQ  = dS K          (A = dS K-major,  B = K MN-major)
// P/Pd/dS live in shared memory as bf16 in the layout the forward kernel uses for P (one thread per
// query row writes its row). Query rows are processed in two tiles (rows 0..127, 128..159); rows and
// keys >= L are zero so they contribute nothing to the K-dimension sums.
// ------------------------------------------------------------------------------------------
constexpr int BT_THREADS = 544;                 // warps 0-15: four threads per query row (column quarters); warp 16: TMA + MMA issue
constexpr int BT_ROWT = 512;
constexpr int BT_CTRL_WARP = BT_ROWT / 32;
constexpr int BT_NK = 160;                      // padded key count (UMMA N of S / dP)
constexpr int BT_OP_BYTES = 16384 + 4096;       // one operand: rows 0..127 then rows 128..159
constexpr int BT_SM_Q = 0;
constexpr int BT_SM_K = BT_OP_BYTES;
constexpr int BT_SM_V = 2 * BT_OP_BYTES;
constexpr int BT_SM_DO = 3 * BT_OP_BYTES;
constexpr int BT_SM_P = 4 * BT_OP_BYTES;        // 3 chunks of 64 keys x 128 rows x 128 B  (81920: 1024-aligned)
constexpr int BT_SM_DS = BT_SM_P + 3 * 16384;
constexpr int BT_SM_END = BT_SM_DS + 3 * 16384 + 16384;  // +16 KB: M-side over-read of key tile 1 stays in bounds
constexpr int BT_SM_O = BT_SM_END;               // forward output tile (delta = rowsum(dO * O) straight from smem)
constexpr int BT_SM_BAR = BT_SM_O + BT_OP_BYTES;
constexpr int BT_SMEM_TOTAL = BT_SM_BAR + 256 + 1024;
constexpr int BT_TM_S = 0, BT_TM_DQ = 192, BT_TM_DK = 256, BT_TM_DV = 384;

#ifdef A2V_ATTN_TRACE
__device__ long long g_bt_trace[64];
#define BT_TR(k) do { if (blockIdx.x == 0 && it == 2) g_bt_trace[k] = clock64(); } while (0)
#else
#define BT_TR(k) do { } while (0)
#endif

template <bool DROP>
__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmQ32,
                        const __grid_constant__ CUtensorMap tmG128, const __grid_constant__ CUtensorMap tmG32,
                        const __grid_constant__ CUtensorMap tmO128, const __grid_constant__ CUtensorMap tmO32,
                        const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BT_SM_BAR);
    uint64_t* bar_load = bars;      // TMA -> everyone
    uint64_t* bar_s = bars + 1;     // S ready
    uint64_t* bar_p = bars + 2;     // P written (128 arrivals)
    uint64_t* bar_dp = bars + 3;    // dP ready
    uint64_t* bar_ds = bars + 4;    // dS written (128 arrivals)
    uint64_t* bar_dq = bars + 5;    // dQ of this tile ready
    uint64_t* bar_epi = bars + 6;   // row threads done with the dK / dV accumulators of this head
    uint64_t* bar_free = bars + 7;  // every MMA of a row tile retired: P / dS buffers (and, after the last tile, the operands) are free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    __shared__ __align__(16) float s_pos[BT_NK];
    __shared__ float s_delta[4][BT_NK];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, D = p.D;
    const int n_mt = L > 128 ? 2 : 1;
    const int n_kt = n_mt;
    const int heads_total = p.batch * p.H;
    // Work distribution. With the fused qkv-bias gradient (p.dbias) every CTA keeps ONE head index h for its whole life
    // (CTA c: h = c % H, sequences b = c / H, c / H + n_h, ...), so the column sums of its 3 x 64 dqkv columns
    // accumulate in 192 shared-memory floats and reach global memory once per CTA instead of once per row.
    int first_head = blockIdx.x, head_step = gridDim.x;
    if (p.dbias != nullptr) {
        const int hf = blockIdx.x % p.H;
        const int n_h = ((int)gridDim.x - hf + p.H - 1) / p.H;  // CTAs that share this head index
        first_head = (blockIdx.x / p.H) * p.H + hf;
        head_step = n_h * p.H;
    }
    __shared__ float s_bias[3 * HD];
    if (tid < 3 * HD) s_bias[tid] = 0.f;

    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, BT_ROWT);
        mbar_init(bar_dp, 1);
        mbar_init(bar_ds, BT_ROWT);
        mbar_init(bar_dq, 1);
        mbar_init(bar_epi, BT_ROWT);
        mbar_init(bar_free, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == BT_CTRL_WARP) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    if (warp == BT_CTRL_WARP) {
        // ================================================================ control warp
        // Issue order per row tile m (the tensor pipe retires in order):
        //   S_m | dP_m | dV += Pd_m^T dO_m | S_{m+1} | dQ_m | dK += dS_m^T Q_m
        // S of the NEXT tile goes ahead of dQ/dK of this one, dQ is committed before the dK products so its
        // rows are stored while they run, and the operands of the next head are fetched as soon as the last
        // product of this head has retired -- before the row threads have drained dK / dV.
        if (elect_one()) {
            tma_prefetch_desc(&tmQ128);
            tma_prefetch_desc(&tmQ32);
            tma_prefetch_desc(&tmG128);
            tma_prefetch_desc(&tmG32);
            tma_prefetch_desc(&tmO128);
            tma_prefetch_desc(&tmO32);
            const uint32_t id_s = umma_idesc_bf16(128, BT_NK, false, false);
            const uint32_t id_dq = umma_idesc_bf16(128, 64, false, true);
            const uint32_t id_t = umma_idesc_bf16(128, 64, true, true);
            // descriptor bases; a byte offset advances the 14-bit (address >> 4) field (smem < 256 KB: no carry)
            const uint64_t dK_kmaj = umma_smem_desc(sbase + BT_SM_K, 0, 1024);
            const uint64_t dV_kmaj = umma_smem_desc(sbase + BT_SM_V, 0, 1024);
            const uint64_t dQ_base = umma_smem_desc(sbase + BT_SM_Q, 0, 1024);
            const uint64_t dG_base = umma_smem_desc(sbase + BT_SM_DO, 0, 1024);
            const uint64_t dP_mn = umma_smem_desc(sbase + BT_SM_P, 16384, 1024);
            const uint64_t dDS_mn = umma_smem_desc(sbase + BT_SM_DS, 16384, 1024);
            const uint64_t dDS_k = umma_smem_desc(sbase + BT_SM_DS, 0, 1024);
#define BT_ADV(desc, bytes) ((desc) + (uint64_t)((uint32_t)(bytes) >> 4))
            const int ksteps = (L + 15) >> 4;  // 16-key steps of dQ = dS K over the keys actually loaded
            uint32_t it = 0;       // heads processed by this CTA
            uint32_t ph_tile = 0;  // parity of the per-tile barriers (bar_p, bar_ds, bar_free)
            auto issue_s = [&](int m) {
                const uint64_t qd = BT_ADV(dQ_base, m * 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + BT_TM_S, BT_ADV(qd, k * 32), BT_ADV(dK_kmaj, k * 32), id_s, k > 0 ? 1u : 0u);
                umma_commit(bar_s);
            };
            auto load_head = [&](int head) {
                const int b = head / p.H, h = head - b * p.H;
                const uint32_t bytes = (uint32_t)(5 * 16384 + (n_mt > 1 ? 5 * 4096 : 0));
                mbar_expect_tx(bar_load, bytes);
                tma_load_3d(smem + BT_SM_Q, &tmQ128, bar_load, h * HD, 0, b);
                tma_load_3d(smem + BT_SM_K, &tmQ128, bar_load, D + h * HD, 0, b);
                tma_load_3d(smem + BT_SM_V, &tmQ128, bar_load, 2 * D + h * HD, 0, b);
                tma_load_3d(smem + BT_SM_DO, &tmG128, bar_load, h * HD, 0, b);
                tma_load_3d(smem + BT_SM_O, &tmO128, bar_load, h * HD, 0, b);
                if (n_mt > 1) {
                    tma_load_3d(smem + BT_SM_Q + 16384, &tmQ32, bar_load, h * HD, 128, b);
                    tma_load_3d(smem + BT_SM_K + 16384, &tmQ32, bar_load, D + h * HD, 128, b);
                    tma_load_3d(smem + BT_SM_V + 16384, &tmQ32, bar_load, 2 * D + h * HD, 128, b);
                    tma_load_3d(smem + BT_SM_DO + 16384, &tmG32, bar_load, h * HD, 128, b);
                    tma_load_3d(smem + BT_SM_O + 16384, &tmO32, bar_load, h * HD, 128, b);
                }
            };
            if (first_head < heads_total) load_head(first_head);
            for (int head = first_head; head < heads_total; head += head_step, ++it) {
                BT_TR(0);
                mbar_wait_sleep(bar_load, it & 1);
                tc_fence_after();
                BT_TR(2);
                issue_s(0);
                BT_TR(3);
                for (int m = 0; m < n_mt; ++m) {
                    const uint64_t qd = BT_ADV(dQ_base, m * 16384), gd = BT_ADV(dG_base, m * 16384);
                    const int krows = m == 0 ? 8 : 2;  // 16-row K steps over the query rows of this tile
                    // P written -> dP_m = dO_m V^T (the TMEM columns of S), dV += Pd_m^T dO_m
                    mbar_wait_sleep(bar_p, ph_tile);
                    tc_fence_after();
                    BT_TR(4 + 5 * m);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + BT_TM_S, BT_ADV(gd, k * 32), BT_ADV(dV_kmaj, k * 32), id_s, k > 0 ? 1u : 0u);
                    umma_commit(bar_dp);
                    if (m == 0 && it > 0) {  // the previous head's dK / dV accumulators must have been read out
                        mbar_wait_sleep(bar_epi, (it - 1) & 1);
                        tc_fence_after();
                    }
                    for (int kt = 0; kt < n_kt; ++kt) {
                        const uint64_t pa = BT_ADV(dP_mn, kt * 32768);
#pragma unroll 2
                        for (int k = 0; k < krows; ++k)
                            umma_bf16(tmem_base + BT_TM_DV + kt * 64, BT_ADV(pa, k * 2048), BT_ADV(gd, k * 2048), id_t,
                                      (m > 0 || k > 0) ? 1u : 0u);
                    }
                    BT_TR(5 + 5 * m);
                    // dS written -> S_{m+1}, dQ_m = dS_m K, dK += dS_m^T Q_m
                    mbar_wait_sleep(bar_ds, ph_tile);
                    tc_fence_after();
                    BT_TR(6 + 5 * m);
                    if (m + 1 < n_mt) issue_s(m + 1);
                    // K extent = keys actually loaded (zero-filled up to the next multiple of 16): never multiply
                    // a zero dS column with stale shared memory
#pragma unroll 2
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_bf16(tmem_base + BT_TM_DQ, BT_ADV(dDS_k, (ks >> 2) * 16384 + (ks & 3) * 32),
                                  BT_ADV(dK_kmaj, ks * 2048), id_dq, ks > 0 ? 1u : 0u);
                    umma_commit(bar_dq);
                    for (int kt = 0; kt < n_kt; ++kt) {
                        const uint64_t da = BT_ADV(dDS_mn, kt * 32768);
#pragma unroll 2
                        for (int k = 0; k < krows; ++k)
                            umma_bf16(tmem_base + BT_TM_DK + kt * 64, BT_ADV(da, k * 2048), BT_ADV(qd, k * 2048), id_t,
                                      (m > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_free);
                    BT_TR(7 + 5 * m);
                    if (m == n_mt - 1) {  // operands free once every product of this head has retired
                        mbar_wait_sleep(bar_free, ph_tile);
                        if (head + head_step < heads_total) load_head(head + head_step);
                        BT_TR(13);
                    }
                    ph_tile ^= 1;
                }
            }
#undef BT_ADV
        }
        __syncwarp();
    } else {
        // ================================================================ row threads
        // 512 threads: four per query row, each owning a quarter of the key columns in 16-column units
        // (3 + 3 + 2 + 2 of the 10 units), so every phase between two barriers is short.
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int rt = tid & 127;        // row inside the 128-row tile
        const int qtr = tid >> 7;        // column quarter
        const int u_lo = qtr == 0 ? 0 : (qtr == 1 ? 3 : (qtr == 2 ? 6 : 8));
        const int u_hi = qtr == 0 ? 3 : (qtr == 1 ? 6 : (qtr == 2 ? 8 : 10));
        const float inv_keep = DROP ? 1.0f / (1.0f - p.drop_p) : 1.0f;
        const uint32_t drop_thr = attn_drop_threshold(p.drop_p);
        const bf16* dout = reinterpret_cast<const bf16*>(p.dout);
        const bf16* outp = reinterpret_cast<const bf16*>(p.out);
        bf16* dqkv = reinterpret_cast<bf16*>(p.dqkv);
        const float scale2 = p.sm_scale * LOG2E;
        uint32_t it = 0, ph_s = 0, ph_dp = 0, ph_dq = 0;
        uint32_t tiles_done = 0;  // row tiles finished by this CTA (bar_free completes once per tile)
        // token position (thread j < 160) and row lse of the NEXT head are fetched one head ahead
        float nx_pos = 0.f, nx_lse0 = 0.f, nx_lse1 = 0.f;
        auto prefetch_head = [&](int head) {
            if (head >= heads_total) return;
            const int b = head / p.H;
            if (tid < BT_NK) nx_pos = tid < L ? (float)(p.pos != nullptr ? p.pos[(long long)b * L + tid] : tid) : 0.f;
            nx_lse0 = rt < L ? p.lse[(long long)head * L + rt] * LOG2E : 0.f;
            nx_lse1 = (n_mt > 1 && 128 + rt < L) ? p.lse[(long long)head * L + 128 + rt] * LOG2E : 0.f;
        };
        prefetch_head(first_head);
        for (int head = first_head; head < heads_total; head += head_step, ++it) {
            const int b = head / p.H, h = head - b * p.H;
            const long long bh = head;
            const float coef = head_coef(p, h);
            const float coef2 = coef * LOG2E;
            float dc_part = 0.f;
            if (tid == 0) BT_TR(16);
            if (tid < BT_NK) s_pos[tid] = nx_pos;
            const float lse_m[2] = {nx_lse0, nx_lse1};
            prefetch_head(head + head_step);
            // delta_i = dO_i . O_i of both row tiles from the TMA-staged tiles (each thread: its 16 of the 64 dims)
            mbar_wait_sleep(bar_load, it & 1);
            float delta_m[2] = {0.f, 0.f};
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (m < n_mt && (m == 0 || rt < 32)) {
                    const uint8_t* gt = smem + BT_SM_DO + m * 16384 + rt * 128;
                    const uint8_t* ot = smem + BT_SM_O + m * 16384 + rt * 128;
                    float d = 0.f;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int unit = (qtr * 2 + k) ^ (rt & 7);
                        const uint4 a = *reinterpret_cast<const uint4*>(gt + unit * 16);
                        const uint4 c = *reinterpret_cast<const uint4*>(ot + unit * 16);
                        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z),
                                     a3 = unpack_bf16x2(a.w);
                        const float2 c0 = unpack_bf16x2(c.x), c1 = unpack_bf16x2(c.y), c2 = unpack_bf16x2(c.z),
                                     c3 = unpack_bf16x2(c.w);
                        d += a0.x * c0.x + a0.y * c0.y + a1.x * c1.x + a1.y * c1.y + a2.x * c2.x + a2.y * c2.y +
                             a3.x * c3.x + a3.y * c3.y;
                    }
                    delta_m[m] = d;  // quarter of the dot product; the four quarters meet in shared memory
                }
            }
            s_delta[qtr][rt] = delta_m[0];
            if (rt < 32) s_delta[qtr][128 + rt] = delta_m[1];
            asm volatile("bar.sync 1, 512;" ::: "memory");
            delta_m[0] = s_delta[0][rt] + s_delta[1][rt] + s_delta[2][rt] + s_delta[3][rt];
            {
                const int r1 = 128 + (rt & 31);
                delta_m[1] = s_delta[0][r1] + s_delta[1][r1] + s_delta[2][r1] + s_delta[3][r1];
            }
            if (tid == 0) BT_TR(17);
            for (int m = 0; m < n_mt; ++m) {
                const int i = m * 128 + rt;           // query row
                const bool row_ok = i < L;
                const bool row_used = m == 0 || rt < 32;   // rows inside the K extent of this tile
                const float delta = delta_m[m], lse_i = lse_m[m];
                const float fpos_i = row_ok ? s_pos[i] : 0.f;
                uint8_t* prow = smem + BT_SM_P + rt * 128;
                uint8_t* drow = smem + BT_SM_DS + rt * 128;
                uint32_t keepbits[3] = {0xffffu, 0xffffu, 0xffffu};  // dropout keep flags of this thread's <= 3 units

                // ---- P (undropped, parked in the dS buffer) and Pd (dropout applied) from S
                if (tid == 0) BT_TR(18 + 8 * m);
                mbar_wait_sleep(bar_s, ph_s);
                ph_s ^= 1;
                tc_fence_after();
                if (tid == 0) BT_TR(19 + 8 * m);
                if (tiles_done > 0) mbar_wait_sleep(bar_free, (tiles_done - 1) & 1);  // P / dS buffers of the previous tile
                if (row_used) {
                    const uint32_t row_key = DROP ? attn_row_key(p.seed, bh, L, i) : 0u;
                    // rows beyond L: lse = +inf makes every probability exactly 0 without a per-element test
                    const float nlse = row_ok ? -lse_i : -INFINITY;
#pragma unroll 1
                    for (int uu = 0; uu < 3; ++uu) {
                        const int u = u_lo + uu;
                        if (u >= u_hi) break;
                        uint32_t raw[16];
                        tmem_ld_32x16(tmem_base + lane_off + BT_TM_S + u * 16, raw);
                        tmem_ld_wait();
                        const bool full_unit = u * 16 + 16 <= L;  // warp-uniform: no key of this unit is padding
                        uint32_t kb = 0;
#pragma unroll
                        for (int g8 = 0; g8 < 2; ++g8) {
                            float pv[8], pd[8];
#pragma unroll
                            for (int g4 = 0; g4 < 2; ++g4) {
                                const int j0 = u * 16 + g8 * 8 + g4 * 4;
                                const float4 pj = *reinterpret_cast<const float4*>(&s_pos[j0]);
                                const float pjs[4] = {pj.x, pj.y, pj.z, pj.w};
                                uint32_t kf[4] = {1u, 1u, 1u, 1u};
                                if (DROP) {
                                    const uint2 bits = attn_bits4(row_key, j0 >> 2);
                                    kf[0] = (bits.x & 0xffffu) >= drop_thr; kf[1] = (bits.x >> 16) >= drop_thr;
                                    kf[2] = (bits.y & 0xffffu) >= drop_thr; kf[3] = (bits.y >> 16) >= drop_thr;
                                }
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float t = fmaf(__uint_as_float(raw[g8 * 8 + g4 * 4 + e]), scale2, nlse);
                                    float pr = ex2_approx(fmaf(fabsf(fpos_i - pjs[e]), -coef2, t));
                                    if (!full_unit && j0 + e >= L) pr = 0.f;
                                    pv[g4 * 4 + e] = pr;
                                    pd[g4 * 4 + e] = kf[e] ? pr * inv_keep : 0.f;
                                    if (DROP) kb |= kf[e] << (g8 * 8 + g4 * 4 + e);
                                }
                            }
                            uint4 v, w;
                            v.x = pack_bf16x2(pd[0], pd[1]); v.y = pack_bf16x2(pd[2], pd[3]);
                            v.z = pack_bf16x2(pd[4], pd[5]); v.w = pack_bf16x2(pd[6], pd[7]);
                            w.x = pack_bf16x2(pv[0], pv[1]); w.y = pack_bf16x2(pv[2], pv[3]);
                            w.z = pack_bf16x2(pv[4], pv[5]); w.w = pack_bf16x2(pv[6], pv[7]);
                            const int unit = ((u & 3) * 2 + g8) ^ (rt & 7);
                            *reinterpret_cast<uint4*>(prow + (u >> 2) * 16384 + unit * 16) = v;
                            *reinterpret_cast<uint4*>(drow + (u >> 2) * 16384 + unit * 16) = w;
                        }
                        if (uu == 0) keepbits[0] = kb;
                        else if (uu == 1) keepbits[1] = kb;
                        else keepbits[2] = kb;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_p);
                if (tid == 0) BT_TR(20 + 8 * m);

                // ---- dS = P * (dP * keep - delta), d(alibi scale)
                mbar_wait_sleep(bar_dp, ph_dp);
                ph_dp ^= 1;
                tc_fence_after();
                if (tid == 0) BT_TR(21 + 8 * m);
                if (row_used) {
#pragma unroll 1
                    for (int uu = 0; uu < 3; ++uu) {
                        const int u = u_lo + uu;
                        if (u >= u_hi) break;
                        uint32_t raw[16];
                        tmem_ld_32x16(tmem_base + lane_off + BT_TM_S + u * 16, raw);
                        tmem_ld_wait();
                        const uint32_t kb = uu == 0 ? keepbits[0] : (uu == 1 ? keepbits[1] : keepbits[2]);
#pragma unroll
                        for (int g8 = 0; g8 < 2; ++g8) {
                            const int unit = ((u & 3) * 2 + g8) ^ (rt & 7);
                            uint4* slot = reinterpret_cast<uint4*>(drow + (u >> 2) * 16384 + unit * 16);
                            const uint4 pw = *slot;
                            const float2 p0 = unpack_bf16x2(pw.x), p1 = unpack_bf16x2(pw.y), p2 = unpack_bf16x2(pw.z),
                                         p3 = unpack_bf16x2(pw.w);
                            const float pr[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
                            float ds[8];
                            const float4 pa = *reinterpret_cast<const float4*>(&s_pos[u * 16 + g8 * 8]);
                            const float4 pb = *reinterpret_cast<const float4*>(&s_pos[u * 16 + g8 * 8 + 4]);
                            const float pjs[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float dp = __uint_as_float(raw[g8 * 8 + e]);
                                if (DROP) dp = ((kb >> (g8 * 8 + e)) & 1u) ? dp * inv_keep : 0.f;
                                // P is exactly 0 for rows / keys beyond L, so dS is too
                                ds[e] = pr[e] * (dp - delta);
                                dc_part = fmaf(-ds[e], fabsf(fpos_i - pjs[e]), dc_part);
                            }
                            uint4 v;
                            v.x = pack_bf16x2(ds[0], ds[1]); v.y = pack_bf16x2(ds[2], ds[3]);
                            v.z = pack_bf16x2(ds[4], ds[5]); v.w = pack_bf16x2(ds[6], ds[7]);
                            *slot = v;
                        }
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_ds);
                if (tid == 0) BT_TR(22 + 8 * m);

                // ---- dQ rows of this tile
                mbar_wait_sleep(bar_dq, ph_dq);
                ph_dq ^= 1;
                tc_fence_after();
                if (tid == 0) BT_TR(23 + 8 * m);
                if (row_used) {
                    uint32_t raw[16];
                    tmem_ld_32x16(tmem_base + lane_off + BT_TM_DQ + qtr * 16, raw);
                    tmem_ld_wait();
                    if (row_ok) {
                        bf16* o = dqkv + ((long long)b * L + i) * 3 * D + h * HD + qtr * 16;
#pragma unroll
                        for (int d = 0; d < 16; d += 8) {
                            uint4 v;
                            v.x = pack_bf16x2(__uint_as_float(raw[d]) * p.sm_scale, __uint_as_float(raw[d + 1]) * p.sm_scale);
                            v.y = pack_bf16x2(__uint_as_float(raw[d + 2]) * p.sm_scale, __uint_as_float(raw[d + 3]) * p.sm_scale);
                            v.z = pack_bf16x2(__uint_as_float(raw[d + 4]) * p.sm_scale, __uint_as_float(raw[d + 5]) * p.sm_scale);
                            v.w = pack_bf16x2(__uint_as_float(raw[d + 6]) * p.sm_scale, __uint_as_float(raw[d + 7]) * p.sm_scale);
                            *reinterpret_cast<uint4*>(o + d) = v;
                        }
                    }
                    if (p.dbias != nullptr) {  // column sums of this warp's 32 rows x 16 dQ columns
                        float cv[16];
#pragma unroll
                        for (int d = 0; d < 16; ++d) cv[d] = row_ok ? __uint_as_float(raw[d]) * p.sm_scale : 0.f;
                        const float cs = warp_colsum<16>(cv, lane);
                        if (lane < 16) atomicAdd(&s_bias[qtr * 16 + lane], cs);
                    }
                }
                tc_fence_before();
                if (tid == 0) BT_TR(24 + 8 * m);
                ++tiles_done;
            }
            // ---- dK, dV rows: quarter -> (dK | dV, key tile); final once the last tile's products retired
            mbar_wait_sleep(bar_free, (tiles_done - 1) & 1);
            tc_fence_after();
            if (tid == 0) BT_TR(39);
            {
                const int which = qtr & 1, kt = qtr >> 1;
                const int j = kt * 128 + rt;
                if (kt < n_kt && (kt == 0 || rt < 32)) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t raw[32];
                        tmem_ld_32x32(tmem_base + lane_off + (which == 0 ? BT_TM_DK : BT_TM_DV) + kt * 64 + c * 32, raw);
                        tmem_ld_wait();
                        if (j < L) {
                            const float sc = which == 0 ? p.sm_scale : 1.0f;
                            bf16* o = dqkv + ((long long)b * L + j) * 3 * D + (which + 1) * D + h * HD + c * 32;
#pragma unroll
                            for (int d = 0; d < 32; d += 8) {
                                uint4 v;
                                v.x = pack_bf16x2(__uint_as_float(raw[d]) * sc, __uint_as_float(raw[d + 1]) * sc);
                                v.y = pack_bf16x2(__uint_as_float(raw[d + 2]) * sc, __uint_as_float(raw[d + 3]) * sc);
                                v.z = pack_bf16x2(__uint_as_float(raw[d + 4]) * sc, __uint_as_float(raw[d + 5]) * sc);
                                v.w = pack_bf16x2(__uint_as_float(raw[d + 6]) * sc, __uint_as_float(raw[d + 7]) * sc);
                                *reinterpret_cast<uint4*>(o + d) = v;
                            }
                        }
                        if (p.dbias != nullptr) {  // column sums of this warp's 32 key rows x 32 dK / dV columns
                            const float sc = which == 0 ? p.sm_scale : 1.0f;
                            float cv[32];
#pragma unroll
                            for (int d = 0; d < 32; ++d) cv[d] = j < L ? __uint_as_float(raw[d]) * sc : 0.f;
                            const float cs = warp_colsum<32>(cv, lane);
                            atomicAdd(&s_bias[(which + 1) * HD + c * 32 + lane], cs);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_epi);
            if (tid == 0) BT_TR(40);
            // d(alibi_scale[h]) += slope_h * sum(-dS * dist)  (only where the clamped scale is active)
            dc_part = warp_sum(dc_part);
            if (lane == 0 && p.dalibi_scale != nullptr && p.alibi_scale != nullptr && p.slopes != nullptr &&
                p.alibi_scale[h * p.alibi_scale_stride] >= 0.f)
                atomicAdd(p.dalibi_scale + h * p.alibi_scale_stride, dc_part * p.slopes[h]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.dbias != nullptr && tid < 3 * HD && first_head < heads_total) {
        const int hf = blockIdx.x % p.H;
        atomicAdd(p.dbias + (tid / HD) * D + hf * HD + (tid % HD), s_bias[tid]);
    }
    if (warp == BT_CTRL_WARP) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// fp32 validation-mode kernels (CUDA cores, one thread per query row)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_fwd_ref_kernel(const AttnParams p) {
    const int L = p.L, D = p.D, h = blockIdx.y, b = blockIdx.z;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= L) return;
    const float* qkv = reinterpret_cast<const float*>(p.qkv);
    const long long bh = (long long)b * p.H + h;
    const float coef = head_coef(p, h);
    const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    float q[HD], o[HD];
    const float* qrow = qkv + ((long long)b * L + i) * 3 * D + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        q[d] = qrow[d] * p.sm_scale;
        o[d] = 0.f;
    }
    const int pi = p.pos != nullptr ? p.pos[(long long)b * L + i] : i;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < L; ++j) {
        const float* krow = qkv + ((long long)b * L + j) * 3 * D + D + h * HD;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s += q[d] * krow[d];
        const int pj = p.pos != nullptr ? p.pos[(long long)b * L + j] : j;
        s -= coef * fabsf((float)(pi - pj));
        const float mn = fmaxf(m, s);
        const float al = __expf(m - mn);
        float e = __expf(s - mn);
        l = l * al + e;
        if (p.drop_p > 0.f) e = attn_keep(p.seed, bh, L, i, j, p.drop_p) ? e * inv_keep : 0.f;
        const float* vrow = krow + D;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = o[d] * al + e * vrow[d];
        m = mn;
    }
    float* orow = reinterpret_cast<float*>(p.out) + ((long long)b * L + i) * D + h * HD;
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) orow[d] = o[d] * inv;
    if (p.lse != nullptr) p.lse[bh * L + i] = m + __logf(l);
}

// dqkv must be zero-initialised (dK / dV are accumulated with atomics)
__global__ void __launch_bounds__(128) attn_bwd_ref_kernel(const AttnParams p) {
    __shared__ float s_dc[4];
    const int L = p.L, D = p.D, h = blockIdx.y, b = blockIdx.z;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const float* qkv = reinterpret_cast<const float*>(p.qkv);
    float* dqkv = reinterpret_cast<float*>(p.dqkv);
    const long long bh = (long long)b * p.H + h;
    const float coef = head_coef(p, h);
    const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    float dc = 0.f;
    if (i < L) {
        float q[HD], go[HD], dq[HD];
        const float* qrow = qkv + ((long long)b * L + i) * 3 * D + h * HD;
        const float* gorow = reinterpret_cast<const float*>(p.dout) + ((long long)b * L + i) * D + h * HD;
        const float* orow = reinterpret_cast<const float*>(p.out) + ((long long)b * L + i) * D + h * HD;
        float delta = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = qrow[d];
            go[d] = gorow[d];
            dq[d] = 0.f;
            delta += go[d] * orow[d];
        }
        const float lse = p.lse[bh * L + i];
        const int pi = p.pos != nullptr ? p.pos[(long long)b * L + i] : i;
        for (int j = 0; j < L; ++j) {
            const float* krow = qkv + ((long long)b * L + j) * 3 * D + D + h * HD;
            const float* vrow = krow + D;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                s += q[d] * krow[d];
                dp += go[d] * vrow[d];
            }
            const int pj = p.pos != nullptr ? p.pos[(long long)b * L + j] : j;
            const float dist = fabsf((float)(pi - pj));
            const float pr = __expf(s * p.sm_scale - coef * dist - lse);
            float ks = 1.f;
            if (p.drop_p > 0.f) ks = attn_keep(p.seed, bh, L, i, j, p.drop_p) ? inv_keep : 0.f;
            const float ds = pr * (dp * ks - delta);
            dc -= ds * dist;
            float* dk = dqkv + ((long long)b * L + j) * 3 * D + D + h * HD;
            float* dv = dk + D;
            const float pk = pr * ks;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                dq[d] += ds * krow[d];
                atomicAdd(dk + d, ds * q[d] * p.sm_scale);
                atomicAdd(dv + d, pk * go[d]);
            }
        }
        float* dqrow = dqkv + ((long long)b * L + i) * 3 * D + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) dqrow[d] = dq[d] * p.sm_scale;
    }
    dc = warp_sum(dc);
    if ((threadIdx.x & 31) == 0) s_dc[threadIdx.x >> 5] = dc;
    __syncthreads();
    if (threadIdx.x == 0 && p.dalibi_scale != nullptr && p.alibi_scale != nullptr && p.slopes != nullptr) {
        const float s = s_dc[0] + s_dc[1] + s_dc[2] + s_dc[3];
        if (p.alibi_scale[h * p.alibi_scale_stride] >= 0.f)
            atomicAdd(p.dalibi_scale + h * p.alibi_scale_stride, s * p.slopes[h]);
    }
}

EncodeTiledFn2 attn_tensor_map_encoder() {
    static EncodeTiledFn2 encode = nullptr;  // idempotent initialisation: a race only repeats the lookup
    if (encode == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym)
            return nullptr;
        encode = reinterpret_cast<EncodeTiledFn2>(sym);
    }
    return encode;
}

int attn_make_map(CUtensorMap* out, const void* base, int cols, int L, int batch, int box_rows) {
    EncodeTiledFn2 encode = attn_tensor_map_encoder();
    if (encode == nullptr) {
        a2v_set_error("attention: cuTensorMapEncodeTiled not available");
        return A2V_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)L * cols * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        a2v_set_error("attention: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return A2V_ERR_CUDA;
    }
    return A2V_OK;
}

int validate_attn(const a2v_attn_desc* d, AttnParams& p) {
    A2V_REQUIRE(d != nullptr, "attention: NULL descriptor");
    A2V_REQUIRE(d->dtype == A2V_F32 || d->dtype == A2V_BF16, "attention: bad dtype");
    A2V_REQUIRE(d->qkv && d->out, "attention: NULL qkv/out");
    A2V_REQUIRE(d->batch > 0 && d->L > 0 && d->H > 0 && d->batch <= 65535 && d->H <= 65535,
                "attention: bad extents batch=%d L=%d H=%d", d->batch, d->L, d->H);
    A2V_REQUIRE(d->head_dim == HD, "attention: only head_dim 64 is supported (got %d)", d->head_dim);
    A2V_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, "attention: bad dropout probability");
    p.qkv = d->qkv; p.out = d->out; p.lse = d->lse; p.pos = d->pos;
    p.slopes = d->slopes; p.alibi_scale = d->alibi_scale; p.alibi_scale_stride = d->alibi_scale_stride;
    p.batch = d->batch; p.L = d->L; p.H = d->H; p.D = d->H * HD;
    p.sm_scale = d->sm_scale; p.drop_p = d->drop_p; p.seed = d->seed;
    p.dout = d->dout; p.dqkv = d->dqkv; p.dalibi_scale = d->dalibi_scale;
    p.qk_bound = d->qk_bound;
    p.delta = nullptr;
    p.dq_acc = nullptr;
    p.dbias = nullptr;
    p.head_filter = 0;
    return A2V_OK;
}

}  // namespace a2v

using namespace a2v;

#ifdef A2V_ATTN_TRACE
extern "C" int a2v_debug_attn_trace(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_bt_trace, sizeof(long long) * (size_t)(n < 64 ? n : 64));
}
extern "C" int a2v_debug_attn_fwd_trace(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_ft_trace, sizeof(long long) * (size_t)(n < 64 ? n : 64));
}
#endif

extern "C" int a2v_attn_qk_bound(const void* qkv_bf16, float* bound, int batch, int L, int H, a2v_stream_t stream) {
    A2V_REQUIRE(qkv_bf16 && bound && batch > 0 && L > 0 && H > 0 && batch <= 65535, "attn_qk_bound: bad arguments");
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(qkv_bf16) & 15) == 0, "attn_qk_bound: qkv not 16-byte aligned");
    attn_qk_bound_kernel<<<dim3(H, batch, ceil_div(L, QKB_CHUNK)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(qkv_bf16), bound, L, H);
    return a2v_check_launch("attn_qk_bound");
}

extern "C" int a2v_attn_fwd(const a2v_attn_desc* d, a2v_stream_t stream) {
    AttnParams p;
    int rc = validate_attn(d, p);
    if (rc != A2V_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid(ceil_div(p.L, 128), p.H, p.batch);
    if (d->dtype == A2V_F32) {
        attn_fwd_ref_kernel<<<grid, 128, 0, st>>>(p);
        return a2v_check_launch("attn_fwd_ref");
    }
    // short sequences (the student's kept tokens): persistent single-pass kernel; A2V_ATTN_SHORT=0 keeps the flash kernel
    static int use_short = -1;
    if (use_short < 0) {
        const char* e = getenv("A2V_ATTN_SHORT");
        use_short = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    if (use_short == 1 && p.L <= ATTN_SHORT_LMAX) return attn_fwd_short_launch(p, st);
    // contiguous sequences with the q/k bound: single-pass stream kernel first, the flash kernel below then only runs the
    // heads whose norms are too large for a fixed reference exponent (A2V_ATTN_STREAM=0: flash kernel for everything)
    static int use_stream = -1;
    if (use_stream < 0) {
        const char* e = getenv("A2V_ATTN_STREAM");
        use_stream = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    if (use_stream == 1 && p.pos == nullptr && p.qk_bound != nullptr && p.L > ATTN_SHORT_LMAX) {
        rc = attn_fwd_stream_launch(p, st);
        if (rc != A2V_OK) return rc;
        p.head_filter = 1;
        grid.x = 1;  // one CTA per (head, sequence): exits at once unless the stream kernel declined the head
    }
    EncodeTiledFn2 encode = attn_tensor_map_encoder();
    if (encode == nullptr) {
        a2v_set_error("attention: cuTensorMapEncodeTiled not available");
        return A2V_ERR_CUDA;
    }
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0, "attention: qkv not 16-byte aligned");
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)(3 * p.D), (cuuint64_t)p.L, (cuuint64_t)p.batch};
    cuuint64_t strides[2] = {(cuuint64_t)(3 * p.D) * 2, (cuuint64_t)p.L * 3 * p.D * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.qkv), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        a2v_set_error("attention: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return A2V_ERR_CUDA;
    }
    const bool has_pos = p.pos != nullptr, drop = p.drop_p > 0.f;
    const bool trim = p.L <= 512;
#define A2V_ATT_GO(P_, D_)                                                                                         \
    do {                                                                                                           \
        const void* kf = trim ? reinterpret_cast<const void*>(attn_fwd_tcgen05_kernel<P_, D_, true>)               \
                              : reinterpret_cast<const void*>(attn_fwd_tcgen05_kernel<P_, D_, false>);             \
        if (a2v_ensure_dynamic_smem(kf, ATT_SMEM_TOTAL) != A2V_OK) return A2V_ERR_CUDA;                            \
        if (trim) attn_fwd_tcgen05_kernel<P_, D_, true><<<grid, 128, ATT_SMEM_TOTAL, st>>>(tm, p);                 \
        else attn_fwd_tcgen05_kernel<P_, D_, false><<<grid, 128, ATT_SMEM_TOTAL, st>>>(tm, p);                     \
    } while (0)
    if (has_pos && drop) A2V_ATT_GO(true, true);
    else if (has_pos) A2V_ATT_GO(true, false);
    else if (drop) A2V_ATT_GO(false, true);
    else A2V_ATT_GO(false, false);
#undef A2V_ATT_GO
    return a2v_check_launch("attn_fwd_tcgen05");
}

extern "C" int a2v_attn_bwd(const a2v_attn_desc* d, a2v_stream_t stream) {
    AttnParams p;
    int rc = validate_attn(d, p);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->dout && d->dqkv && d->lse, "attention backward: dout / dqkv / lse are required");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d->dtype == A2V_F32) {
        dim3 grid(ceil_div(p.L, 128), p.H, p.batch);
        attn_bwd_ref_kernel<<<grid, 128, 0, st>>>(p);
        return a2v_check_launch("attn_bwd_ref");
    }
    A2V_REQUIRE(d->bwd_algo >= 0 && d->bwd_algo <= 2, "attention backward: bwd_algo must be 0 (auto), 1 (resident) or 2 (tiled)");
    if (d->bwd_algo == 2 || (d->bwd_algo == 0 && p.L > BWD_LMAX)) {
        // any length: key tiles resident, query tiles streamed (attention_bwd.cu); needs the prepared workspace
        A2V_REQUIRE(d->workspace != nullptr &&
                        (size_t)d->workspace_bytes >= a2v_attn_bwd_workspace_bytes(p.batch, p.L, p.H),
                    "attention backward (bf16, tiled): workspace of a2v_attn_bwd_workspace_bytes() bytes required");
        A2V_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.dout) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0,
                    "attention backward: qkv / dout / workspace not 16-byte aligned");
        const size_t rows = (size_t)p.batch * (size_t)p.L;
        p.delta = reinterpret_cast<const float*>(d->workspace);
        p.dq_acc = reinterpret_cast<float*>(d->workspace) + ((rows * (size_t)p.H + 3) & ~(size_t)3);
        return attn_bwd_tiled_launch(p, st);
    }
    A2V_REQUIRE(p.L <= BWD_LMAX,
                "attention backward (bf16): the shared-memory-resident kernel supports at most %d tokens per "
                "sequence, got %d", BWD_LMAX, p.L);
    if (d->dqkv_colsum != nullptr) {
        A2V_REQUIRE(p.H <= a2v_num_sms(), "attention backward: fused qkv-bias gradient needs H <= number of SMs");
        p.dbias = d->dqkv_colsum;
    }
    EncodeTiledFn2 encode = attn_tensor_map_encoder();
    if (encode == nullptr) {
        a2v_set_error("attention: cuTensorMapEncodeTiled not available");
        return A2V_ERR_CUDA;
    }
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.dout) & 15) == 0,
                "attention backward: qkv / dout not 16-byte aligned");
    A2V_REQUIRE(d->out != nullptr && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0,
                "attention backward: forward output missing or not 16-byte aligned");
    CUtensorMap maps[6];  // qkv (128 / 32-row boxes), dout, out
    for (int i = 0; i < 6; ++i) {
        const bool is_q = i < 2;
        const int cols = is_q ? 3 * p.D : p.D;
        cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)p.L, (cuuint64_t)p.batch};
        cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)p.L * cols * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)((i & 1) ? 32 : 128), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                            const_cast<void*>(is_q ? p.qkv : (i < 4 ? p.dout : (const void*)p.out)), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            a2v_set_error("attention backward: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return A2V_ERR_CUDA;
        }
    }
    if (a2v_ensure_dynamic_smem(p.drop_p > 0.f ? reinterpret_cast<const void*>(attn_bwd_tcgen05_kernel<true>)
                                               : reinterpret_cast<const void*>(attn_bwd_tcgen05_kernel<false>),
                                BT_SMEM_TOTAL) != A2V_OK)
        return A2V_ERR_CUDA;
    const int heads = p.batch * p.H;
    const int grid = heads < a2v_num_sms() ? heads : a2v_num_sms();
    if (p.drop_p > 0.f)
        attn_bwd_tcgen05_kernel<true><<<grid, BT_THREADS, BT_SMEM_TOTAL, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
    else
        attn_bwd_tcgen05_kernel<false><<<grid, BT_THREADS, BT_SMEM_TOTAL, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
    return a2v_check_launch("attn_bwd_tcgen05");
}
