// Attention forward for the student's short sequences (L <= 160 kept tokens per clone; nn/modalities/modules.py:368-410
// on the rows nn/modalities/base.py:427-455 keeps): persistent, single pass.
//
// With ~148 keys the whole key range fits ONE tcgen05 product (S = Q K^T, N = 160), so there is no online softmax, no
// running-maximum rescale and no second key tile: per work item (clone b, head h, query tile) a row thread reads its
// 160 scores twice out of tensor memory (pass 1: row maximum, pass 2: exponentials), writes the probabilities back INTO
// the score columns as packed bf16 (the TMEM A operand of P.V) and the control warp issues one P.V product.
// Two threads share a query row (80 key columns each, partial maximum / sum exchanged through shared memory): with one
// thread per row a scheduler partition holds one or two runnable warps and the dependent FMA / MUFU chains leave half the
// issue slots empty (measured). Persistent CTAs (two per SM: 224 TMEM columns, 100 KB shared memory) walk the work items with the K/V tiles double
// buffered; S of the next item is issued the moment P.V of this one retires, so it runs under this item's epilogue.
// The 20-odd remainder rows of a sequence (query tile 1) occupy the TMEM lanes of a warp that ROTATES with the head
// index, so the four scheduler partitions share that work instead of all of it landing on warp 0.
// The general flash kernel of attention.cu serves every longer sequence; both produce the same out / lse contract.
#include "attention_common.cuh"

namespace a2v {

constexpr int SH_ROWT = 256;             // warps 0-7: TWO threads per query row (80 key columns each)
constexpr int SH_CTRL_WARP = SH_ROWT / 32;
constexpr int SH_THREADS = SH_ROWT + 32;  // warp 8: TMA + MMA issue
constexpr int SH_NK = ATTN_SHORT_LMAX;   // key extent of S (UMMA N)
constexpr int SH_KV_BYTES = 16384 + 4096;  // 128 + 32 rows of 128 B
constexpr int SH_SM_Q = 0;                               // single buffer (free again as soon as S has been computed)
constexpr int SH_SM_K = 16384;                           // 2 stages
constexpr int SH_SM_V = SH_SM_K + 2 * SH_KV_BYTES;       // 2 stages
constexpr int SH_SM_POS = SH_SM_V + 2 * SH_KV_BYTES;     // 2 x 160 floats: coef * key position
constexpr int SH_SM_COEF = SH_SM_POS + 2 * SH_NK * 4;    // 64 floats: ALiBi coefficient (log2 units) per head
constexpr int SH_SM_X = SH_SM_COEF + 64 * 4;             // 2 (max, sum) x 2 (item parity) x 2 (half) x 128 floats
constexpr int SH_SM_BAR = SH_SM_X + 8 * 128 * 4;
constexpr int SH_SMEM_TOTAL = SH_SM_BAR + 128 + 1024;
constexpr int SH_TM_S = 0, SH_TM_O = 160, SH_TM_P1 = 224;  // P of keys 0..95 over score columns 0..47, keys 96..159 in the spare 32
constexpr int SH_SPLIT = 96;                                // keys [0, 96) -> thread half 0, [96, 160) -> half 1

#ifdef A2V_ATTN_TRACE
__device__ long long g_sh_trace[128];
// items 6..9 of CTA 0: row thread 0 (slots 0..7 per item) and the control thread (slots 8..15 per item)
#define SH_TR(k) do { if (blockIdx.x == 0 && i >= 6u && i < 10u && (tid == 0 || tid == SH_ROWT)) g_sh_trace[(i - 6u) * 16 + (k)] = clock64(); } while (0)
#else
#define SH_TR(k) do { } while (0)
#endif

struct ShortItem {
    int b, h, qt, rot;  // rot: TMEM-lane quarter (warp) that holds the remainder rows of query tile 1
};
__device__ __forceinline__ ShortItem short_item(int item, int n_qt, int H) {
    ShortItem it;
    const int bh = item / n_qt;
    it.qt = item - bh * n_qt;
    it.b = bh / H;
    it.h = bh - it.b * H;
    it.rot = it.qt == 0 ? 0 : (bh & 3);
    return it;
}

template <bool HAS_POS, bool DROP>
__global__ void __launch_bounds__(SH_THREADS, 2)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm32,
                      const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SH_SM_BAR);
    uint64_t* bar_q = bars;       // Q tile landed (one completion per item)
    uint64_t* bar_kv = bars + 1;  // [2] K / V of a stage landed
    uint64_t* bar_s = bars + 3;   // S ready
    uint64_t* bar_p = bars + 4;   // P written (128 arrivals)
    uint64_t* bar_o = bars + 5;   // P.V retired: O ready, S / P columns and the K/V stage free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    float* spos = reinterpret_cast<float*>(smem + SH_SM_POS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, D = p.D, H = p.H;
    const int n_qt = L > 128 ? 2 : 1;
    const int items = p.batch * H * n_qt;
    const int ksteps = (L + 15) >> 4;  // P.V K extent: keys actually present, rounded up to the MMA K of 16

    if (tid == 0) {
        mbar_init(bar_q, 1);
        mbar_init(&bar_kv[0], 1);
        mbar_init(&bar_kv[1], 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, SH_ROWT);
        mbar_init(bar_o, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    float* scoef = reinterpret_cast<float*>(smem + SH_SM_COEF);
    float* sx = reinterpret_cast<float*>(smem + SH_SM_X);
    if (tid < 64) scoef[tid] = tid < H ? head_coef(p, tid) * LOG2E : 0.f;
    if (warp == SH_CTRL_WARP) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == SH_CTRL_WARP) {
        // ================================================================ control warp
        if (elect_one()) {
            tma_prefetch_desc(&tm128);
            tma_prefetch_desc(&tm32);
            const uint32_t idesc_s = umma_idesc_bf16(128, SH_NK, false, false);
            const uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);
            const uint32_t qa = smem_u32(smem + SH_SM_Q);
            auto load_q = [&](int item) {
                const ShortItem w = short_item(item, n_qt, H);
                if (w.qt == 0) {
                    mbar_expect_tx(bar_q, 16384);
                    tma_load_3d(smem + SH_SM_Q, &tm128, bar_q, w.h * HD, 0, w.b);
                } else {  // remainder rows 128..159 -> tile rows 32 * rot .. (the other rows keep stale data: their S / O
                          // rows are never read)
                    mbar_expect_tx(bar_q, 4096);
                    tma_load_3d(smem + SH_SM_Q + w.rot * 4096, &tm32, bar_q, w.h * HD, 128, w.b);
                }
            };
            auto load_kv = [&](int item, int stage) {
                const ShortItem w = short_item(item, n_qt, H);
                uint8_t* ks = smem + SH_SM_K + stage * SH_KV_BYTES;
                uint8_t* vs = smem + SH_SM_V + stage * SH_KV_BYTES;
                mbar_expect_tx(&bar_kv[stage], (uint32_t)(2 * (L > 128 ? SH_KV_BYTES : 16384)));
                tma_load_3d(ks, &tm128, &bar_kv[stage], D + w.h * HD, 0, w.b);
                tma_load_3d(vs, &tm128, &bar_kv[stage], 2 * D + w.h * HD, 0, w.b);
                if (L > 128) {
                    tma_load_3d(ks + 16384, &tm32, &bar_kv[stage], D + w.h * HD, 128, w.b);
                    tma_load_3d(vs + 16384, &tm32, &bar_kv[stage], 2 * D + w.h * HD, 128, w.b);
                }
            };
            auto issue_s = [&](int stage) {
                const uint32_t ka = smem_u32(smem + SH_SM_K + stage * SH_KV_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem_base + SH_TM_S, umma_smem_desc(qa + k * 32, 0, 1024), umma_smem_desc(ka + k * 32, 0, 1024),
                              idesc_s, k > 0 ? 1u : 0u);
                umma_commit(bar_s);
            };
            const int first = blockIdx.x, step = gridDim.x;
            if (first < items) {
                load_q(first);
                load_kv(first, 0);
                if (first + step < items) load_kv(first + step, 1);
            }
            uint32_t i = 0;
            for (int item = first; item < items; item += step, ++i) {
                const int s = (int)(i & 1u);
                if (i == 0) {
                    mbar_wait(bar_q, 0);
                    mbar_wait(&bar_kv[0], 0);
                    tc_fence_after();
                    issue_s(0);
                }
                // S(i) computed: the Q buffer is free for the next item's tile
                mbar_wait_sleep(bar_s, i & 1u);
                SH_TR(8);
                if (item + step < items) load_q(item + step);
                // P(i) written -> O = P V
                mbar_wait_sleep(bar_p, i & 1u);
                SH_TR(9);
                tc_fence_after();
                {
                    const uint32_t va = smem_u32(smem + SH_SM_V + s * SH_KV_BYTES);
                    for (int k = 0; k < ksteps; ++k)  // 16 keys = 8 packed P columns per step (keys >= 96: the spare columns)
                        umma_bf16_ts(tmem_base + SH_TM_O, tmem_base + (k < 6 ? SH_TM_S + k * 8 : SH_TM_P1 + (k - 6) * 8),
                                     umma_smem_desc(va + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
                    umma_commit(bar_o);
                    SH_TR(10);
                }
                // P.V retired: score columns free -> S of the next item runs under this item's epilogue
                mbar_wait_sleep(bar_o, i & 1u);
                SH_TR(11);
                if (item + step < items) {
                    mbar_wait(bar_q, (i + 1u) & 1u);
                    mbar_wait(&bar_kv[s ^ 1], ((i + 1u) >> 1) & 1u);
                    tc_fence_after();
                    issue_s(s ^ 1);
                    SH_TR(12);
                    if (item + 2 * step < items) load_kv(item + 2 * step, s);
                }
            }
        }
        __syncwarp();
    } else {
        // ================================================================ row threads
        const int qw = warp & 3;     // TMEM lane quarter of this warp
        const int half = warp >> 2;  // key-column half: keys [80 * half, 80 * half + 80)
        const int rt = qw * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qw * 32) << 16;
        const uint32_t ts = tmem_base + lane_off + SH_TM_S;
        const int c0 = half * SH_SPLIT;                 // first key of this thread
        const int nch = half == 0 ? SH_SPLIT / 16 : (SH_NK - SH_SPLIT) / 16;  // its 16-key chunks (6 / 4)
        // P destination: the thread of half 0 writes over score columns it has already consumed ([8c, 8c+8) lies below
        // [0, 16(c+1))), the thread of half 1 into the 32 spare TMEM columns -- no thread ever overwrites scores its
        // partner still has to read, so the probabilities stream out chunk by chunk without a second barrier
        const uint32_t tp = tmem_base + lane_off + (half == 0 ? SH_TM_S : SH_TM_P1);
        const float scale2 = p.sm_scale * LOG2E;
        const float inv_keep = DROP ? 1.0f / (1.0f - p.drop_p) : 1.0f;
        const uint32_t drop_thr = attn_drop_threshold(p.drop_p);
        const int kext = ksteps * 16;  // keys covered by P.V
        // token positions of the NEXT item are fetched one item ahead (a global load on the critical path of every item
        // costs more than the whole softmax of a remainder tile)
        int nx = 0;
        auto prefetch_pos = [&](int item) {
            if (!HAS_POS || item >= items || tid >= SH_NK) return;
            const ShortItem w = short_item(item, n_qt, H);
            nx = tid < L ? p.pos[(long long)w.b * L + tid] : 0;
        };
        prefetch_pos(blockIdx.x);
        uint32_t i = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++i) {
            const ShortItem w = short_item(item, n_qt, H);
            SH_TR(0);
            const long long bh = (long long)w.b * H + w.h;
            // kp[j] = coef2 * position of key j (double buffered: the previous item's readers may still be in pass 2)
            float* kp = spos + (i & 1u) * SH_NK;
            float* xm = sx + (i & 1u) * 256;        // [half][row]: partial row maxima
            float* xl = sx + 512 + (i & 1u) * 256;  // [half][row]: partial row sums
            if (tid < SH_NK) kp[tid] = scoef[w.h] * (float)(HAS_POS ? nx : tid);
            prefetch_pos(item + gridDim.x);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            SH_TR(1);
            // this thread's query row
            int qi;
            bool active;
            if (w.qt == 0) {
                qi = rt;
                active = qi < L;
            } else {
                qi = 128 + lane;
                active = qw == w.rot && qi < L;
            }
            const bool warp_active = w.qt == 0 ? (qw * 32 < L) : (qw == w.rot);
            const float cpi = active ? kp[qi] : 0.f;

            mbar_wait_sleep(bar_s, i & 1u);
            SH_TR(2);
            tc_fence_after();
            float m_row = -INFINITY, l_row = 0.f;
            if (warp_active) {
                // ---- pass 1: partial row maximum of  scale2 * s - |cp_i - cp_j|  over this thread's existing keys
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    const int k0 = c0 + c * 16;
                    if (k0 >= L) break;
                    uint32_t raw[16];
                    tmem_ld_32x16(ts + k0, raw);
                    tmem_ld_wait();
                    const bool ragged = k0 + 16 > L;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 k4 = *reinterpret_cast<const float4*>(kp + k0 + q * 4);
                        const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float v = fmaf(__uint_as_float(raw[q * 4 + e]), scale2, -fabsf(cpi - kk[e]));
                            if (ragged && k0 + q * 4 + e >= L) v = -INFINITY;
                            m_row = fmaxf(m_row, v);
                        }
                    }
                }
                xm[half * 128 + rt] = m_row;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(2 + qw) : "memory");  // the two warps that share these 32 rows
            SH_TR(3);
            if (warp_active) {
                m_row = fmaxf(m_row, xm[(half ^ 1) * 128 + rt]);
                // ---- pass 2: exponentials, partial row sum, dropout, P (packed bf16) streamed to tensor memory
                const uint32_t row_key = DROP ? attn_row_key(p.seed, bh, L, qi) : 0u;
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    const int k0 = c0 + c * 16;
                    if (k0 >= kext) break;
                    uint32_t raw[16];
                    tmem_ld_32x16(ts + k0, raw);
                    tmem_ld_wait();
                    const bool ragged = k0 + 16 > L;
                    uint32_t pk[8];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float e8[8];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float4 k4 = *reinterpret_cast<const float4*>(kp + k0 + u * 8 + q * 4);
                            const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float t = fmaf(__uint_as_float(raw[u * 8 + q * 4 + e]), scale2, -m_row);
                                float ex = ex2_approx(t - fabsf(cpi - kk[e]));
                                if (ragged && k0 + u * 8 + q * 4 + e >= L) ex = 0.f;
                                e8[q * 4 + e] = ex;
                                l_row += ex;
                            }
                        }
                        if (DROP) {
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                const uint2 bits = attn_bits4(row_key, (k0 + u * 8) / 4 + g);
                                e8[4 * g + 0] = (bits.x & 0xffffu) >= drop_thr ? e8[4 * g + 0] * inv_keep : 0.f;
                                e8[4 * g + 1] = (bits.x >> 16) >= drop_thr ? e8[4 * g + 1] * inv_keep : 0.f;
                                e8[4 * g + 2] = (bits.y & 0xffffu) >= drop_thr ? e8[4 * g + 2] * inv_keep : 0.f;
                                e8[4 * g + 3] = (bits.y >> 16) >= drop_thr ? e8[4 * g + 3] * inv_keep : 0.f;
                            }
                        }
                        pk[u * 4 + 0] = pack_bf16x2(e8[0], e8[1]);
                        pk[u * 4 + 1] = pack_bf16x2(e8[2], e8[3]);
                        pk[u * 4 + 2] = pack_bf16x2(e8[4], e8[5]);
                        pk[u * 4 + 3] = pack_bf16x2(e8[6], e8[7]);
                    }
                    tmem_st_32x8(tp + c * 8, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                }
                tmem_st_wait();
                xl[half * 128 + rt] = l_row;  // read by the partner after bar_o (the mbarrier orders it)
            }
            SH_TR(4);
            tc_fence_before();
            mbar_arrive(bar_p);

            // epilogue: O / l -> out (32 of the 64 columns per thread), log-sum-exp
            if (warp_active) {
                mbar_wait_sleep(bar_o, i & 1u);
                SH_TR(5);
                tc_fence_after();
                l_row += xl[(half ^ 1) * 128 + rt];
                const float inv_l = active ? 1.0f / l_row : 0.f;
                bf16* orow = reinterpret_cast<bf16*>(p.out) + ((long long)w.b * L + (active ? qi : 0)) * D + w.h * HD + half * 32;
                uint32_t r0[32];
                tmem_ld_32x32(tmem_base + lane_off + SH_TM_O + half * 32, r0);
                tmem_ld_wait();
                if (active) {
#pragma unroll
                    for (int d = 0; d < 32; d += 8) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(r0[d]) * inv_l, __uint_as_float(r0[d + 1]) * inv_l);
                        v.y = pack_bf16x2(__uint_as_float(r0[d + 2]) * inv_l, __uint_as_float(r0[d + 3]) * inv_l);
                        v.z = pack_bf16x2(__uint_as_float(r0[d + 4]) * inv_l, __uint_as_float(r0[d + 5]) * inv_l);
                        v.w = pack_bf16x2(__uint_as_float(r0[d + 6]) * inv_l, __uint_as_float(r0[d + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + d) = v;
                    }
                    if (half == 0 && p.lse != nullptr) p.lse[bh * L + qi] = (m_row + log2f(l_row)) * LN2;
                }
                SH_TR(6);
                tc_fence_before();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SH_CTRL_WARP) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

int attn_fwd_short_launch(const AttnParams& p, cudaStream_t st) {
    A2V_REQUIRE(p.L <= ATTN_SHORT_LMAX, "attention forward (short): at most %d tokens", ATTN_SHORT_LMAX);
    A2V_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0,
                "attention forward: qkv / out not 16-byte aligned");
    CUtensorMap t128, t32;
    int rc = attn_make_map(&t128, p.qkv, 3 * p.D, p.L, p.batch, 128);
    if (rc != A2V_OK) return rc;
    rc = attn_make_map(&t32, p.qkv, 3 * p.D, p.L, p.batch, 32);
    if (rc != A2V_OK) return rc;
    const bool has_pos = p.pos != nullptr, drop = p.drop_p > 0.f;
    const int n_qt = p.L > 128 ? 2 : 1;
    const long long items = (long long)p.batch * p.H * n_qt;
    const int cap = 2 * a2v_num_sms();
    const int grid = items < cap ? (int)items : cap;
#define A2V_SH_GO(P_, D_)                                                                                              \
    do {                                                                                                               \
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_short_kernel<P_, D_>), SH_SMEM_TOTAL) != A2V_OK) \
            return A2V_ERR_CUDA;                                                                                       \
        attn_fwd_short_kernel<P_, D_><<<grid, SH_THREADS, SH_SMEM_TOTAL, st>>>(t128, t32, p);                          \
    } while (0)
    if (has_pos && drop) A2V_SH_GO(true, true);
    else if (has_pos) A2V_SH_GO(true, false);
    else if (drop) A2V_SH_GO(false, true);
    else A2V_SH_GO(false, false);
#undef A2V_SH_GO
    return a2v_check_launch("attn_fwd_short");
}

}  // namespace a2v

#ifdef A2V_ATTN_TRACE
extern "C" int a2v_debug_attn_short_trace(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, a2v::g_sh_trace, sizeof(long long) * (size_t)(n < 128 ? n : 128));
}
#endif
