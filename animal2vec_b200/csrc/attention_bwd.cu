// Attention backward for ANY sequence length (bf16, tcgen05): the finetune path (full T = 2000 frames with
// gradients, nn/wav2vec2.py:437-444 -> nn/modalities/modules.py:368-410 under autograd) and the 48 kHz
// pretraining configuration (Tk ~ 852 kept tokens). The shared-memory-resident kernel of attention.cu keeps serving
// the short student sequences (L <= 160).
//
// Work item = (batch b, head h, key tile j of 128 keys), persistent CTAs. K_j / V_j stay in shared memory while the
// query tiles i stream through a two-stage TMA ring (Q_i, dO_i); per (i, j) pair five products run on tcgen05 with TMEM
// accumulators, exactly the algebra of the resident kernel:
//   S  = Q_i K_j^T            -> P = exp2(S * scale + alibi - lse_i)           (row threads, TMEM -> smem bf16)
//   dP = dO_i V_j^T           -> dS = P * (dP * keep - delta_i)
//   dV_j += Pd^T dO_i,  dK_j += dS^T Q_i      (TMEM accumulators that live across the query tiles)
//   dQ_i  = dS K_j            -> fp32 vector atomics into a (batch, L, D) accumulator (one contribution per key tile)
// Contiguous sequences with an ALiBi slope visit only the query tiles inside the same 2^-50 window the forward uses
// (a2v_attn_qk_bound): probabilities outside it are below fp32 resolution of the row sum, so are their gradients.
// a2v_attn_bwd_prepare computes delta = rowsum(dO * O) and clears the dQ accumulator, a2v_attn_bwd_finish scales and
// rounds it into dqkv: one kernel launch per C-ABI call, like every other entry point.
#include "attention_common.cuh"

namespace a2v {

constexpr int GB_THREADS = 544;  // warps 0-15: four threads per query row (32 key columns each); warp 16: TMA + MMA issue
constexpr int GB_ROWT = 512;
constexpr int GB_CTRL_WARP = GB_ROWT / 32;
constexpr int GB_SM_K = 0;
constexpr int GB_SM_V = 16384;
constexpr int GB_SM_Q = 32768;    // 2 stages
constexpr int GB_SM_DO = 65536;   // 2 stages
constexpr int GB_SM_P = 98304;    // 2 chunks of 64 keys x 128 rows x 128 B
constexpr int GB_SM_DS = 131072;  // same layout
constexpr int GB_SM_BAR = 163840;
constexpr int GB_SMEM_TOTAL = GB_SM_BAR + 256 + 1024;
constexpr int GB_TM_S = 0, GB_TM_DQ = 128, GB_TM_DK = 192, GB_TM_DV = 256;
constexpr float GB_SKIP_LOG2 = 50.0f;  // same window as the forward (ATT_SKIP_LOG2)

// query-tile range [i0, i1) visited for key tile kt of head (b, h); symmetric to the forward's key-tile window
__device__ __forceinline__ void gb_tile_range(const AttnParams& p, int b, int h, int kt, int n_t, int& i0, int& i1) {
    i0 = 0;
    i1 = n_t;
    if (p.pos == nullptr && p.qk_bound != nullptr) {
        const float c2 = head_coef(p, h) * LOG2E;
        if (c2 > 0.f) {
            const float2 b2 = reinterpret_cast<const float2*>(p.qk_bound)[b * p.H + h];
            const float qk = sqrtf(b2.x) * sqrtf(b2.y) * 1.002f;
            const float w = (2.f * qk * (p.sm_scale * LOG2E) + GB_SKIP_LOG2) / c2;
            if (w < 1.0e6f) {
                const int wi = (int)w + 1;
                const int k0 = kt * 128;
                const int lo = k0 - 127 - wi;
                i0 = lo < 0 ? 0 : lo / 128 + 1;
                const int hi = (k0 + 126 + wi) / 128 + 1;
                i1 = hi < n_t ? hi : n_t;
            }
        }
    }
}

template <bool DROP>
__global__ void __launch_bounds__(GB_THREADS, 1)
attn_bwd_tiled_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GB_SM_BAR);
    uint64_t* bar_kv = bars;        // K_j / V_j landed (once per item)
    uint64_t* bar_q = bars + 1;     // [2] Q_i / dO_i of a stage landed
    uint64_t* bar_s = bars + 3;     // S ready
    uint64_t* bar_p = bars + 4;     // P written (512 arrivals)
    uint64_t* bar_dp = bars + 5;    // dP ready
    uint64_t* bar_ds = bars + 6;    // dS written (512 arrivals)
    uint64_t* bar_dq = bars + 7;    // dQ of this pair ready
    uint64_t* bar_free = bars + 8;  // every MMA of this pair retired: P / dS buffers and the Q / dO stage are free
    uint64_t* bar_epi = bars + 9;   // row threads drained dK / dV of this item (512 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    __shared__ __align__(16) float s_pos[128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, D = p.D, H = p.H;
    const int n_t = (L + 127) >> 7;
    const int items = p.batch * H * n_t;

    if (tid == 0) {
        mbar_init(bar_kv, 1);
        mbar_init(&bar_q[0], 1);
        mbar_init(&bar_q[1], 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, GB_ROWT);
        mbar_init(bar_dp, 1);
        mbar_init(bar_ds, GB_ROWT);
        mbar_init(bar_dq, 1);
        mbar_init(bar_free, 1);
        mbar_init(bar_epi, GB_ROWT);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == GB_CTRL_WARP) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    if (warp == GB_CTRL_WARP) {
        // ================================================================ control warp
        if (elect_one()) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            const uint32_t id_s = umma_idesc_bf16(128, 128, false, false);
            const uint32_t id_dq = umma_idesc_bf16(128, 64, false, true);
            const uint32_t id_t = umma_idesc_bf16(128, 64, true, true);
            const uint64_t dK_base = umma_smem_desc(sbase + GB_SM_K, 0, 1024);
            const uint64_t dV_base = umma_smem_desc(sbase + GB_SM_V, 0, 1024);
            const uint64_t dQ_base = umma_smem_desc(sbase + GB_SM_Q, 0, 1024);
            const uint64_t dG_base = umma_smem_desc(sbase + GB_SM_DO, 0, 1024);
            const uint64_t dP_mn = umma_smem_desc(sbase + GB_SM_P, 16384, 1024);
            const uint64_t dDS_mn = umma_smem_desc(sbase + GB_SM_DS, 16384, 1024);
            const uint64_t dDS_k = umma_smem_desc(sbase + GB_SM_DS, 0, 1024);
#define GB_ADV(desc, bytes) ((desc) + (uint64_t)((uint32_t)(bytes) >> 4))
            auto load_kv = [&](int item) {
                const int kt = item % n_t, bh = item / n_t;
                const int b = bh / H, h = bh - b * H;
                mbar_expect_tx(bar_kv, 32768);
                tma_load_3d(smem + GB_SM_K, &tmQKV, bar_kv, D + h * HD, kt * 128, b);
                tma_load_3d(smem + GB_SM_V, &tmQKV, bar_kv, 2 * D + h * HD, kt * 128, b);
            };
            auto load_q = [&](int b, int h, int i, int stage) {
                mbar_expect_tx(&bar_q[stage], 32768);
                tma_load_3d(smem + GB_SM_Q + stage * 16384, &tmQKV, &bar_q[stage], h * HD, i * 128, b);
                tma_load_3d(smem + GB_SM_DO + stage * 16384, &tmDO, &bar_q[stage], h * HD, i * 128, b);
            };
            auto first_loads = [&](int item, uint32_t gt_next) {  // K/V of `item` and its first two query tiles
                const int kt = item % n_t, bh = item / n_t;
                const int b = bh / H, h = bh - b * H;
                int i0, i1;
                gb_tile_range(p, b, h, kt, n_t, i0, i1);
                load_kv(item);
                load_q(b, h, i0, (int)(gt_next & 1u));
                if (i0 + 1 < i1) load_q(b, h, i0 + 1, (int)((gt_next + 1u) & 1u));
            };
            auto issue_s = [&](int stage) {
                const uint64_t qd = GB_ADV(dQ_base, stage * 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + GB_TM_S, GB_ADV(qd, k * 32), GB_ADV(dK_base, k * 32), id_s, k > 0 ? 1u : 0u);
                umma_commit(bar_s);
            };
            uint32_t it = 0, gt = 0;
            uint32_t qph0 = 0, qph1 = 0;
            auto wait_q = [&](int stage) {
                if (stage == 0) { mbar_wait(&bar_q[0], qph0); qph0 ^= 1u; }
                else { mbar_wait(&bar_q[1], qph1); qph1 ^= 1u; }
                tc_fence_after();
            };
            if ((int)blockIdx.x < items) first_loads(blockIdx.x, 0u);
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                const int kt = item % n_t, bh = item / n_t;
                const int b = bh / H, h = bh - b * H;
                int i0, i1;
                gb_tile_range(p, b, h, kt, n_t, i0, i1);
                const int n = i1 - i0;
                for (int t = 0; t < n; ++t, ++gt) {
                    const int s = (int)(gt & 1u);
                    const uint64_t qd = GB_ADV(dQ_base, s * 16384), gd = GB_ADV(dG_base, s * 16384);
                    if (t == 0) {
                        mbar_wait(bar_kv, it & 1u);
                        wait_q(s);
                        issue_s(s);
                    }
                    // P written -> dP = dO_i V_j^T (over the S columns), dV += Pd^T dO_i
                    mbar_wait(bar_p, gt & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + GB_TM_S, GB_ADV(gd, k * 32), GB_ADV(dV_base, k * 32), id_s, k > 0 ? 1u : 0u);
                    umma_commit(bar_dp);
                    if (t == 0 && it > 0) {  // the previous item's dK / dV accumulators must have been read out
                        mbar_wait(bar_epi, (it - 1u) & 1u);
                        tc_fence_after();
                    }
#pragma unroll 2
                    for (int k = 0; k < 8; ++k)  // K extent: the 128 query rows of this tile, 16 per step
                        umma_bf16(tmem_base + GB_TM_DV, GB_ADV(dP_mn, k * 2048), GB_ADV(gd, k * 2048), id_t,
                                  (t > 0 || k > 0) ? 1u : 0u);
                    // dS written -> S of the next pair, dQ = dS K_j, dK += dS^T Q_i
                    mbar_wait(bar_ds, gt & 1u);
                    tc_fence_after();
                    if (t + 1 < n) {
                        wait_q(s ^ 1);
                        issue_s(s ^ 1);
                    }
#pragma unroll 2
                    for (int ks = 0; ks < 8; ++ks)  // K extent: the 128 keys of this tile
                        umma_bf16(tmem_base + GB_TM_DQ, GB_ADV(dDS_k, (ks >> 2) * 16384 + (ks & 3) * 32),
                                  GB_ADV(dK_base, ks * 2048), id_dq, ks > 0 ? 1u : 0u);
                    umma_commit(bar_dq);
#pragma unroll 2
                    for (int k = 0; k < 8; ++k)
                        umma_bf16(tmem_base + GB_TM_DK, GB_ADV(dDS_mn, k * 2048), GB_ADV(qd, k * 2048), id_t,
                                  (t > 0 || k > 0) ? 1u : 0u);
                    umma_commit(bar_free);
                    if (t + 2 < n) {  // refill this stage once its readers (dV, dK of this pair) have retired
                        mbar_wait(bar_free, gt & 1u);
                        load_q(b, h, i0 + t + 2, s);
                    } else if (t == n - 1) {  // item done: everything in shared memory is free
                        mbar_wait(bar_free, gt & 1u);
                        if (item + (int)gridDim.x < items) first_loads(item + gridDim.x, gt + 1u);
                    }
                }
            }
#undef GB_ADV
        }
        __syncwarp();
    } else {
        // ================================================================ row threads
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int rt = tid & 127;  // row inside the 128-row tile
        const int qtr = tid >> 7;  // column quarter: 16-key units 2*qtr, 2*qtr + 1
        const float inv_keep = DROP ? 1.0f / (1.0f - p.drop_p) : 1.0f;
        const uint32_t drop_thr = attn_drop_threshold(p.drop_p);
        bf16* dqkv = reinterpret_cast<bf16*>(p.dqkv);
        const float scale2 = p.sm_scale * LOG2E;
        uint32_t gt = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int kt = item % n_t, bh_i = item / n_t;
            const int b = bh_i / H, h = bh_i - b * H;
            const long long bh = bh_i;
            const int k0 = kt * 128;
            const int nvalid = L - k0;  // keys of this tile that exist
            int i0, i1;
            gb_tile_range(p, b, h, kt, n_t, i0, i1);
            const int n = i1 - i0;
            const float coef2 = head_coef(p, h) * LOG2E;
            float dc_part = 0.f;
            // the previous item's readers of s_pos all passed its last bar_ds before anyone got here (bar_dq follows it)
            if (tid < 128) {
                const int kj = k0 + tid;
                s_pos[tid] = kj < L ? (float)(p.pos != nullptr ? p.pos[(long long)b * L + kj] : kj) : 0.f;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            for (int t = 0; t < n; ++t, ++gt) {
                const int i = (i0 + t) * 128 + rt;  // query row
                const bool row_ok = i < L;
                float lse_i = 0.f, delta = 0.f, fpos_i = 0.f;
                if (row_ok) {
                    lse_i = p.lse[bh * L + i] * LOG2E;
                    delta = p.delta[bh * L + i];
                    fpos_i = (float)(p.pos != nullptr ? p.pos[(long long)b * L + i] : i);
                }
                uint8_t* prow = smem + GB_SM_P + rt * 128;
                uint8_t* drow = smem + GB_SM_DS + rt * 128;
                uint32_t keepbits[2] = {0xffffu, 0xffffu};

                // ---- P (undropped, parked in the dS buffer) and Pd (dropout applied) from S
                mbar_wait(bar_s, gt & 1u);
                tc_fence_after();
                if (gt > 0) mbar_wait(bar_free, (gt - 1u) & 1u);  // P / dS buffers of the previous pair
                {
                    const uint32_t row_key = DROP ? attn_row_key(p.seed, bh, L, i) : 0u;
                    const float nlse = row_ok ? -lse_i : -INFINITY;  // rows beyond L: every probability exactly 0
#pragma unroll
                    for (int uu = 0; uu < 2; ++uu) {
                        const int u = qtr * 2 + uu;
                        uint32_t raw[16];
                        tmem_ld_32x16(tmem_base + lane_off + GB_TM_S + u * 16, raw);
                        tmem_ld_wait();
                        const bool full_unit = u * 16 + 16 <= nvalid;
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
                                    const uint2 bits = attn_bits4(row_key, (k0 + j0) >> 2);
                                    kf[0] = (bits.x & 0xffffu) >= drop_thr; kf[1] = (bits.x >> 16) >= drop_thr;
                                    kf[2] = (bits.y & 0xffffu) >= drop_thr; kf[3] = (bits.y >> 16) >= drop_thr;
                                }
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float tt = fmaf(__uint_as_float(raw[g8 * 8 + g4 * 4 + e]), scale2, nlse);
                                    float pr = ex2_approx(fmaf(fabsf(fpos_i - pjs[e]), -coef2, tt));
                                    if (!full_unit && j0 + e >= nvalid) pr = 0.f;
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
                        keepbits[uu] = kb;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_p);

                // ---- dS = P * (dP * keep - delta), d(alibi scale)
                mbar_wait(bar_dp, gt & 1u);
                tc_fence_after();
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    const int u = qtr * 2 + uu;
                    uint32_t raw[16];
                    tmem_ld_32x16(tmem_base + lane_off + GB_TM_S + u * 16, raw);
                    tmem_ld_wait();
                    const uint32_t kb = keepbits[uu];
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
                            ds[e] = pr[e] * (dp - delta);  // P is exactly 0 for rows / keys beyond L, so dS is too
                            dc_part = fmaf(-ds[e], fabsf(fpos_i - pjs[e]), dc_part);
                        }
                        uint4 v;
                        v.x = pack_bf16x2(ds[0], ds[1]); v.y = pack_bf16x2(ds[2], ds[3]);
                        v.z = pack_bf16x2(ds[4], ds[5]); v.w = pack_bf16x2(ds[6], ds[7]);
                        *slot = v;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_ds);

                // ---- dQ contribution of this key tile: 16 of the 64 columns per thread, fp32 vector atomics
                mbar_wait(bar_dq, gt & 1u);
                tc_fence_after();
                {
                    uint32_t raw[16];
                    tmem_ld_32x16(tmem_base + lane_off + GB_TM_DQ + qtr * 16, raw);
                    tmem_ld_wait();
                    if (row_ok) {
                        float4* o = reinterpret_cast<float4*>(p.dq_acc + ((long long)b * L + i) * D + h * HD + qtr * 16);
#pragma unroll
                        for (int d4 = 0; d4 < 4; ++d4)
                            atomicAdd(o + d4, make_float4(__uint_as_float(raw[4 * d4]), __uint_as_float(raw[4 * d4 + 1]),
                                                          __uint_as_float(raw[4 * d4 + 2]), __uint_as_float(raw[4 * d4 + 3])));
                    }
                }
                tc_fence_before();
            }
            // ---- dK_j, dV_j rows: quarter -> (dK | dV, column half); final once the last pair's products retired
            mbar_wait(bar_free, (gt - 1u) & 1u);
            tc_fence_after();
            {
                const int which = qtr & 1, half = qtr >> 1;
                const int j = k0 + rt;
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + lane_off + (which == 0 ? GB_TM_DK : GB_TM_DV) + half * 32, raw);
                tmem_ld_wait();
                if (j < L) {
                    const float sc = which == 0 ? p.sm_scale : 1.0f;
                    bf16* o = dqkv + ((long long)b * L + j) * 3 * D + (which + 1) * D + h * HD + half * 32;
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
            }
            tc_fence_before();
            mbar_arrive(bar_epi);
            // d(alibi_scale[h]) += slope_h * sum(-dS * dist)  (only where the clamped scale is active)
            dc_part = warp_sum(dc_part);
            if (lane == 0 && p.dalibi_scale != nullptr && p.alibi_scale != nullptr && p.slopes != nullptr &&
                p.alibi_scale[h * p.alibi_scale_stride] >= 0.f)
                atomicAdd(p.dalibi_scale + h * p.alibi_scale_stride, dc_part * p.slopes[h]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GB_CTRL_WARP) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// delta[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]; the dQ accumulator row (b, i, :) is cleared in the same pass.
// One warp per (b, i) row: 8 lanes per head and 256-column step.
__global__ void __launch_bounds__(256) attn_bwd_prepare_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ out,
                                                               float* __restrict__ delta, float* __restrict__ dq_acc,
                                                               long long rows, int L, int H) {
    const int lane = threadIdx.x & 31;
    const int D = H * HD;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = warp0; r < rows; r += nwarps) {
        const long long b = r / L;
        const int i = (int)(r - b * L);
        for (int c = lane * 8; c < D; c += 256) {
            const uint4 g = *reinterpret_cast<const uint4*>(dout + r * D + c);
            const uint4 o = *reinterpret_cast<const uint4*>(out + r * D + c);
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 a = unpack_bf16x2(gw[k]), c2 = unpack_bf16x2(ow[k]);
                s = fmaf(a.x, c2.x, fmaf(a.y, c2.y, s));
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if ((lane & 7) == 0) delta[(b * H + c / HD) * L + i] = s;
            float4* z = reinterpret_cast<float4*>(dq_acc + r * D + c);
            z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// dqkv[b, i, 0:D] = bf16(dq_acc[b, i, :] * sm_scale)
__global__ void __launch_bounds__(256) attn_bwd_finish_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                              long long rows, int D, float sm_scale) {
    const long long n8 = rows * (D / 8);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / (D / 8);
        const int c = (int)(e - r * (D / 8)) * 8;
        const float4 a = *reinterpret_cast<const float4*>(dq_acc + r * D + c);
        const float4 b = *reinterpret_cast<const float4*>(dq_acc + r * D + c + 4);
        uint4 v;
        v.x = pack_bf16x2(a.x * sm_scale, a.y * sm_scale);
        v.y = pack_bf16x2(a.z * sm_scale, a.w * sm_scale);
        v.z = pack_bf16x2(b.x * sm_scale, b.y * sm_scale);
        v.w = pack_bf16x2(b.z * sm_scale, b.w * sm_scale);
        *reinterpret_cast<uint4*>(dqkv + r * 3 * D + c) = v;
    }
}

// called by a2v_attn_bwd (attention.cu) for bf16 sequences longer than the resident kernel takes
int attn_bwd_tiled_launch(const AttnParams& p, cudaStream_t st) {
    A2V_REQUIRE(p.delta != nullptr && p.dq_acc != nullptr,
                "attention backward (bf16, L > 160): workspace missing -- call a2v_attn_bwd_prepare first and pass the "
                "same workspace (a2v_attn_bwd_workspace_bytes)");
    CUtensorMap tq, tg;
    int rc = attn_make_map(&tq, p.qkv, 3 * p.D, p.L, p.batch, 128);
    if (rc != A2V_OK) return rc;
    rc = attn_make_map(&tg, p.dout, p.D, p.L, p.batch, 128);
    if (rc != A2V_OK) return rc;
    if (a2v_ensure_dynamic_smem(p.drop_p > 0.f ? reinterpret_cast<const void*>(attn_bwd_tiled_kernel<true>)
                                               : reinterpret_cast<const void*>(attn_bwd_tiled_kernel<false>),
                                GB_SMEM_TOTAL) != A2V_OK)
        return A2V_ERR_CUDA;
    const int n_t = (p.L + 127) / 128;
    const long long items = (long long)p.batch * p.H * n_t;
    const int grid = items < a2v_num_sms() ? (int)items : a2v_num_sms();
    if (p.drop_p > 0.f)
        attn_bwd_tiled_kernel<true><<<grid, GB_THREADS, GB_SMEM_TOTAL, st>>>(tq, tg, p);
    else
        attn_bwd_tiled_kernel<false><<<grid, GB_THREADS, GB_SMEM_TOTAL, st>>>(tq, tg, p);
    return a2v_check_launch("attn_bwd_tiled");
}

}  // namespace a2v

using namespace a2v;

// workspace = [delta: batch*L*H floats, rounded up to a multiple of 4][dq_acc: batch*L*H*64 floats]
static inline size_t gb_delta_elems(size_t rows, int H) { return (rows * (size_t)H + 3) & ~(size_t)3; }

extern "C" size_t a2v_attn_bwd_workspace_bytes(int batch, int L, int H) {
    if (batch <= 0 || L <= 0 || H <= 0) return 0;
    const size_t rows = (size_t)batch * (size_t)L;
    return (gb_delta_elems(rows, H) + rows * (size_t)H * HD) * sizeof(float);
}

extern "C" int a2v_attn_bwd_prepare(const a2v_attn_desc* d, a2v_stream_t stream) {
    AttnParams p;
    int rc = validate_attn(d, p);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->dtype == A2V_BF16, "attn_bwd_prepare: bf16 only (the fp32 validation kernels need no workspace)");
    A2V_REQUIRE(d->dout && d->out && d->workspace, "attn_bwd_prepare: dout / out / workspace are required");
    A2V_REQUIRE((size_t)d->workspace_bytes >= a2v_attn_bwd_workspace_bytes(p.batch, p.L, p.H),
                "attn_bwd_prepare: workspace too small");
    A2V_REQUIRE(((reinterpret_cast<uintptr_t>(d->dout) | reinterpret_cast<uintptr_t>(d->out) |
                  reinterpret_cast<uintptr_t>(d->workspace)) & 15) == 0, "attn_bwd_prepare: 16-byte alignment");
    const long long rows = (long long)p.batch * p.L;
    float* delta = reinterpret_cast<float*>(d->workspace);
    float* dq_acc = delta + gb_delta_elems((size_t)rows, p.H);
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)a2v_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    attn_bwd_prepare_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(d->dout), reinterpret_cast<const bf16*>(d->out), delta, dq_acc, rows, p.L, p.H);
    return a2v_check_launch("attn_bwd_prepare");
}

extern "C" int a2v_attn_bwd_finish(const a2v_attn_desc* d, a2v_stream_t stream) {
    AttnParams p;
    int rc = validate_attn(d, p);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->dtype == A2V_BF16 && d->dqkv && d->workspace, "attn_bwd_finish: bf16 dqkv and the workspace are required");
    const long long rows = (long long)p.batch * p.L;
    const float* dq_acc = reinterpret_cast<const float*>(d->workspace) + gb_delta_elems((size_t)rows, p.H);
    long long blocks = (rows * (p.D / 8) + 255) / 256;
    const long long cap = (long long)a2v_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    attn_bwd_finish_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        dq_acc, reinterpret_cast<bf16*>(d->dqkv), rows, p.D, p.sm_scale);
    return a2v_check_launch("attn_bwd_finish");
}
