// Masking / multi-mask cloning data movement (HBM-bound, vectorised, warp-per-row).
//
// Reference behaviour replaced (file:line under /root/reference):
//   nn/modalities/base.py:427-455  make_maskinfo  (argsort x2, gather)  -> a2v_mask_index
//   base.py:244, 457-464           repeat_interleave(M) + zero masking  -> a2v_row_gather (clone map)
//   base.py:278-280, 537-542       x_unmasked + gather(x_pos, ids_keep) -> a2v_row_gather (+add)
//   base.py:162-192                decoder_input: dropout, N(0,std) mask tokens, gather(ids_restore)
//                                                                        -> a2v_row_gather (noise fill, dropout)
// and the autograd backward of each (scatter back / sum over the M clones).
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

// One block per mask row. Kept positions keep their ascending order (the reference's order
// comes from an unstable argsort of a 0/1 tensor and is arbitrary; the *set* is identical).
__global__ void __launch_bounds__(256) mask_index_kernel(const uint8_t* __restrict__ mask, int T, int Tk, int M,
                                                         int* __restrict__ ids_keep, int* __restrict__ ids_restore,
                                                         int* __restrict__ clone_src, int* __restrict__ keep_src_x,
                                                         int* __restrict__ keep_src_clone, int* __restrict__ restore_src,
                                                         int* __restrict__ err) {
    __shared__ int warp_cnt[8];
    __shared__ int running;
    const int r = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const uint8_t* mrow = mask + (long long)r * T;
    for (int t0 = 0; t0 < T; t0 += 256) {
        const int t = t0 + threadIdx.x;
        const bool in = t < T;
        const bool kept = in && (mrow[t] == 0);
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        const int rank_k = before + __popc(bal & ((1u << lane) - 1u));
        if (in) {
            const long long o = (long long)r * T + t;
            if (kept) {
                ids_restore[o] = rank_k;
                if (rank_k < Tk) {
                    const long long ok = (long long)r * Tk + rank_k;
                    ids_keep[ok] = t;
                    keep_src_x[ok] = (r / M) * T + t;
                    keep_src_clone[ok] = r * T + t;
                }
                clone_src[o] = (r / M) * T + t;
                restore_src[o] = rank_k < Tk ? r * Tk + rank_k : -1;
            } else {
                ids_restore[o] = Tk + (t - rank_k);
                clone_src[o] = -1;
                restore_src[o] = -1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
            running += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && running != Tk) atomicExch(err, 1 + r);
}

struct GatherParams {
    const void* src;
    const void* add;
    void* dst;
    const int* idx;
    long long n_dst;
    int D;
    float fill_std;
    unsigned long long fill_seed;
    float drop_p;
    unsigned long long drop_seed;
    int drop_by_src;
};

__device__ __forceinline__ void normal4(unsigned long long seed, unsigned long long idx4, float std, float (&o)[4]) {
    const uint64_t h0 = rng64(seed, 2 * idx4), h1 = rng64(seed, 2 * idx4 + 1);
    const float u0 = fmaxf(u01((uint32_t)h0), 5.96e-8f), u1 = u01((uint32_t)(h0 >> 32));
    const float u2 = fmaxf(u01((uint32_t)h1), 5.96e-8f), u3 = u01((uint32_t)(h1 >> 32));
    const float r0 = sqrtf(-2.f * __logf(u0)) * std, r1 = sqrtf(-2.f * __logf(u2)) * std;
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c);
    o[0] = r0 * c; o[1] = r0 * s;
    __sincosf(6.283185307179586f * u3, &s, &c);
    o[2] = r1 * c; o[3] = r1 * s;
}

// dst[o, :] = (idx[o] >= 0 ? dropout(src[idx[o], :]) : fill) + add[o, :]
template <typename T>
__global__ void __launch_bounds__(256) row_gather_kernel(const GatherParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const T* src = reinterpret_cast<const T*>(p.src);
    const T* add = reinterpret_cast<const T*>(p.add);
    T* dst = reinterpret_cast<T*>(p.dst);
    const int D = p.D;
    const float keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    for (long long o = warp0; o < p.n_dst; o += nwarps) {
        const int s = p.idx[o];
        for (int c = lane * 4; c < D; c += 128) {
            float v[4];
            if (s >= 0) {
                load4(src + (long long)s * D + c, v);
                if (p.drop_p > 0.f) {
                    bool k[4];
                    const unsigned long long e = (unsigned long long)((p.drop_by_src ? (long long)s : o) * D + c) >> 2;
                    drop_keep4(p.drop_seed, e, p.drop_p, k);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = k[j] ? v[j] * keep : 0.f;
                }
            } else if (p.fill_std > 0.f) {
                normal4(p.fill_seed, (unsigned long long)(o * D + c) >> 2, p.fill_std, v);
            } else {
                v[0] = v[1] = v[2] = v[3] = 0.f;
            }
            if (add != nullptr) {
                float a[4];
                load4(add + o * D + c, a);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += a[j];
            }
            store4(dst + o * D + c, v);
        }
    }
}

// dx[b, t, :] = sum over the M clones r = b*M + m that keep t of
//               (d_masked[r, t, :] (optional) + d_unmasked[r, rank(r, t), :] (optional))
template <typename T>
__global__ void __launch_bounds__(256) clone_sum_bwd_kernel(const T* __restrict__ d_masked,
                                                            const T* __restrict__ d_unmasked,
                                                            const int* __restrict__ restore_src, T* __restrict__ dx,
                                                            long long BT, int T_, int M, int D) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long o = warp0; o < BT; o += nwarps) {
        const long long b = o / T_;
        const int t = (int)(o - b * T_);
        for (int c = lane * 4; c < D; c += 128) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int m = 0; m < M; ++m) {
                const long long rt = (b * M + m) * T_ + t;
                const int s = restore_src[rt];
                if (s < 0) continue;
                float v[4];
                if (d_masked != nullptr) {
                    load4(d_masked + rt * D + c, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] += v[j];
                }
                if (d_unmasked != nullptr) {
                    load4(d_unmasked + (long long)s * D + c, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] += v[j];
                }
            }
            store4(dx + o * D + c, acc);
        }
    }
}

__global__ void __launch_bounds__(256) neigh_index_kernel(const int* __restrict__ ids_keep, long long n, int Tk, int T_,
                                                          int taps, int pad, int* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long tok = e / taps;
        const int j = (int)(e - tok * taps);
        const long long r = tok / Tk;
        const int t = ids_keep[tok] + j - pad;
        out[e] = (t >= 0 && t < T_) ? (int)(r * T_ + t) : -1;
    }
}

static int grid_for_rows(long long rows) {
    long long b = ceil_div64(rows, 8);
    long long cap = (long long)a2v_num_sms() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace a2v

using namespace a2v;

extern "C" int a2v_mask_index(const uint8_t* mask, int rows, int T, int Tk, int clones, int32_t* ids_keep,
                              int32_t* ids_restore, int32_t* clone_src, int32_t* keep_src_x, int32_t* keep_src_clone,
                              int32_t* restore_src, int32_t* err_flag, a2v_stream_t stream) {
    A2V_REQUIRE(mask && ids_keep && ids_restore && clone_src && keep_src_x && keep_src_clone && restore_src && err_flag,
                "mask_index: NULL pointer");
    A2V_REQUIRE(rows >= 0 && T > 0 && Tk >= 0 && Tk <= T && clones >= 1, "mask_index: bad extents");
    A2V_REQUIRE((long long)rows * T < (1LL << 31), "mask_index: rows*T must fit int32 row maps");
    if (rows == 0) return A2V_OK;
    mask_index_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mask, T, Tk, clones, ids_keep, ids_restore, clone_src, keep_src_x, keep_src_clone, restore_src, err_flag);
    return a2v_check_launch("mask_index");
}

extern "C" int a2v_neigh_index(const int32_t* ids_keep, int rows, int Tk, int T, int taps, int pad, int32_t* out,
                               a2v_stream_t stream) {
    A2V_REQUIRE(ids_keep && out && rows >= 0 && Tk >= 0 && T > 0 && taps >= 1 && pad >= 0, "neigh_index: bad arguments");
    A2V_REQUIRE((long long)rows * T < (1LL << 31), "neigh_index: rows*T must fit int32 row maps");
    const long long n = (long long)rows * Tk * taps;
    if (n == 0) return A2V_OK;
    long long blocks = ceil_div64(n, 256);
    const long long cap = (long long)a2v_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    neigh_index_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids_keep, n, Tk, T, taps, pad, out);
    return a2v_check_launch("neigh_index");
}

extern "C" int a2v_row_gather(int dtype, const void* src, const int32_t* idx, const void* add, void* dst,
                              int64_t n_dst, int D, float fill_std, uint64_t fill_seed, float drop_p,
                              uint64_t drop_seed, int drop_by_src, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "row_gather: bad dtype");
    A2V_REQUIRE(src && idx && dst, "row_gather: NULL pointer");
    A2V_REQUIRE(D > 0 && D % 4 == 0 && n_dst >= 0, "row_gather: D must be a positive multiple of 4");
    A2V_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "row_gather: bad dropout probability");
    if (n_dst == 0) return A2V_OK;
    GatherParams p{src, add, dst, idx, n_dst, D, fill_std, fill_seed, drop_p, drop_seed, drop_by_src};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == A2V_F32)
        row_gather_kernel<float><<<grid_for_rows(n_dst), 256, 0, st>>>(p);
    else
        row_gather_kernel<bf16><<<grid_for_rows(n_dst), 256, 0, st>>>(p);
    return a2v_check_launch("row_gather");
}

extern "C" int a2v_clone_sum_bwd(int dtype, const void* d_masked, const void* d_unmasked, const int32_t* restore_src,
                                 void* dx, int64_t B, int T, int clones, int D, a2v_stream_t stream) {
    A2V_REQUIRE(dtype == A2V_F32 || dtype == A2V_BF16, "clone_sum_bwd: bad dtype");
    A2V_REQUIRE(restore_src && dx && (d_masked || d_unmasked), "clone_sum_bwd: NULL pointer");
    A2V_REQUIRE(D > 0 && D % 4 == 0 && T > 0 && clones >= 1, "clone_sum_bwd: bad extents");
    if (B == 0) return A2V_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long BT = (long long)B * T;
    if (dtype == A2V_F32)
        clone_sum_bwd_kernel<float><<<grid_for_rows(BT), 256, 0, st>>>(
            (const float*)d_masked, (const float*)d_unmasked, restore_src, (float*)dx, BT, T, clones, D);
    else
        clone_sum_bwd_kernel<bf16><<<grid_for_rows(BT), 256, 0, st>>>(
            (const bf16*)d_masked, (const bf16*)d_unmasked, restore_src, (bf16*)dx, BT, T, clones, D);
    return a2v_check_launch("clone_sum_bwd");
}
