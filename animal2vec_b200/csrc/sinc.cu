// SincNet band-pass front end (fp32 CUDA cores; the filters are regenerated every step).
//
// Reference behaviour replaced (file:line under /root/reference):
//   nn/sinc.py:181-223  _get_sinc_filters  -> a2v_sinc_filters_fwd / _bwd
//   nn/sinc.py:286-313  reflect "same" padding (31/31) and sinc.py:144-151 fp32 F.conv1d
//                       -> a2v_sinc_conv_fwd (padding folded into the tile load, filters in smem)
//   autograd of the conv w.r.t. the filters -> a2v_sinc_conv_wgrad
// Output layout is channels-last (B, N, Cpad) with Cpad = 128 (channel 127 is a zero pad) so
// that the next stage (LayerNorm over channels, then a K = 10*128 GEMM) reads contiguous rows.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/a2v_capi.h"

namespace a2v {

constexpr int SINC_CPAD = 128;
constexpr int SINC_KMAX = 128;

// one block (64 threads) per channel
__global__ void sinc_filters_fwd_kernel(const float* __restrict__ low_hz, const float* __restrict__ band_hz,
                                        const float* __restrict__ n_, const float* __restrict__ window_, int C, int K,
                                        float min_low, float min_band, float nyquist, float* __restrict__ filt) {
    const int c = blockIdx.x;
    const int half = K / 2;
    float* f = filt + (long long)c * K;
    if (c >= C) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) f[i] = 0.f;
        return;
    }
    const float low = min_low + fabsf(low_hz[c]);
    const float high = fminf(fmaxf(low + min_band + fabsf(band_hz[c]), min_low), nyquist);
    const float band = high - low;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float n = n_[i];
        const float left = (sinf(high * n) - sinf(low * n)) / n * 2.f * window_[i];
        const float v = left / (2.f * band);
        f[i] = v;
        f[K - 1 - i] = v;
    }
    if (threadIdx.x == 0) f[half] = (2.f * band) / (2.f * band);
}

__global__ void sinc_filters_bwd_kernel(const float* __restrict__ low_hz, const float* __restrict__ band_hz,
                                        const float* __restrict__ n_, const float* __restrict__ window_, int C, int K,
                                        float min_low, float min_band, float nyquist,
                                        const float* __restrict__ dfilt, float* __restrict__ dlow,
                                        float* __restrict__ dband) {
    const int c = blockIdx.x;
    if (c >= C) return;
    const int half = K / 2;
    const float lraw = low_hz[c], braw = band_hz[c];
    const float low = min_low + fabsf(lraw);
    const float pre = low + min_band + fabsf(braw);
    const float high = fminf(fmaxf(pre, min_low), nyquist);
    const bool unclamped = pre >= min_low && pre <= nyquist;
    const float band = high - low;
    const float* g = dfilt + (long long)c * K;
    float d_high = 0.f, d_low = 0.f;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float n = n_[i], w = window_[i];
        const float G = g[i] + g[K - 1 - i];
        const float gi = (sinf(high * n) - sinf(low * n)) * w / n;  // filter_left * band
        d_high += G * (cosf(high * n) * w / band - gi / (band * band));
        d_low += G * (-cosf(low * n) * w / band + gi / (band * band));
    }
    d_high = warp_sum(d_high);
    d_low = warp_sum(d_low);
    __shared__ float sh[2][2];
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = d_high;
        sh[1][threadIdx.x >> 5] = d_low;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float dh = sh[0][0] + sh[0][1], dl = sh[1][0] + sh[1][1];
        const float sl = lraw > 0.f ? 1.f : (lraw < 0.f ? -1.f : 0.f);
        const float sb = braw > 0.f ? 1.f : (braw < 0.f ? -1.f : 0.f);
        const float dh_eff = unclamped ? dh : 0.f;
        atomicAdd(dlow + c, (dl + dh_eff) * sl);
        atomicAdd(dband + c, dh_eff * sb);
    }
}

__device__ __forceinline__ int reflect_index(int j, int N) {
    if (j < 0) j = -j;
    if (j >= N) j = 2 * (N - 1) - j;
    return j;
}

constexpr int SINC_TT = 128;  // time steps per tile

// y[b, t, c] = sum_k filt[c][k] * x[b, reflect(t + k - K/2)]
template <typename TO>
__global__ void __launch_bounds__(256) sinc_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ filt,
                                                            TO* __restrict__ y, int N, int K) {
    extern __shared__ float sm[];
    float* sf = sm;                      // [K][128]
    float* sx = sm + K * SINC_CPAD;      // [SINC_TT + K - 1 + 3]
    const int b = blockIdx.y, t0 = blockIdx.x * SINC_TT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < K * SINC_CPAD; i += 256) {
        const int k = i / SINC_CPAD, c = i - k * SINC_CPAD;
        sf[i] = filt[(long long)c * K + k];
    }
    const int half = K / 2;
    const float* xb = x + (long long)b * N;
    for (int i = threadIdx.x; i < SINC_TT + K + 2; i += 256) {
        const int j = t0 + i - half;
        sx[i] = (j < N + half) ? xb[reflect_index(j, N)] : 0.f;
    }
    __syncthreads();
    // warp w: time steps w*16 .. w*16+15, 4 at a time; lane: channels lane*4 .. lane*4+3
    for (int g = 0; g < 4; ++g) {
        const int tl = warp * 16 + g * 4;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
        float x0 = sx[tl], x1 = sx[tl + 1], x2 = sx[tl + 2];
        for (int k = 0; k < K; ++k) {
            const float x3 = sx[tl + k + 3];
            const float4 f = *reinterpret_cast<const float4*>(sf + k * SINC_CPAD + lane * 4);
            const float xs[4] = {x0, x1, x2, x3};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                acc[a][0] += xs[a] * f.x;
                acc[a][1] += xs[a] * f.y;
                acc[a][2] += xs[a] * f.z;
                acc[a][3] += xs[a] * f.w;
            }
            x0 = x1; x1 = x2; x2 = x3;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int t = t0 + tl + a;
            if (t < N) store4(y + ((long long)b * N + t) * SINC_CPAD + lane * 4, acc[a]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// The same convolution on tcgen05 (bf16 output): per 128-sample tile the Toeplitz operand
//   A[t][k] = x[b, reflect(t0 + t + k - K/2)]        (128 x 128, K padded with zero filter taps)
// is built in shared memory and multiplied with the (128 channels x 128 taps) filter matrix. Both operands are split
// into bf16 hi + lo pieces and three products accumulate in fp32 (hi*hi + hi*lo + lo*hi: the band-pass sums cancel
// heavily, single bf16 operands would cost ~1e-2 of the output): 24 MMAs of 128 x 128 x 16 per tile, ~1 500 tensor
// clocks against ~16 000 FMA-pipe clocks of the CUDA-core kernel above (which stays for the fp32 validation mode).
// 256 threads: everyone builds the next tile's operand while the tensor pipe works on the current one; thread 0
// issues; warps 0-3 / 4-7 drain columns 0-63 / 64-127 of the double-buffered TMEM accumulator.
// ------------------------------------------------------------------------------------------
constexpr int STC_A_BYTES = 2 * 2 * 16384;     // hi / lo x two 64-tap halves, each a 128-row 128B-swizzled tile
constexpr int STC_SM_B = 0;                    // filters: hi (2 x 16 KB), lo (2 x 16 KB)
constexpr int STC_SM_A = 65536;                // two operand buffers
constexpr int STC_SM_X = STC_SM_A + 2 * STC_A_BYTES;   // fp32 samples of a tile: 128 + 127 (+ pad)
constexpr int STC_SM_BAR = STC_SM_X + 2 * 1024;
constexpr int STC_SMEM_TOTAL = STC_SM_BAR + 64 + 1024;

// element (row r, tap k) of a [128][64] bf16 tile with 128-byte swizzle: 16-byte chunk index XOR (row mod 8)
__device__ __forceinline__ uint32_t stc_chunk_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1) sinc_conv_fwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ filt,
                                                                  bf16* __restrict__ y, int N, int K, int tiles_per_clip,
                                                                  int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + STC_SM_BAR);  // [2]: accumulator buffer i complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = K / 2;

    if (tid == 0) {
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    // filters -> hi / lo bf16 tiles (row = channel, K-major); thread: channel tid & 127, tap half tid >> 7
    const int nsteps = (K + 15) >> 4;  // K steps of 16 taps (taps beyond K are zero: not multiplied at all)
    if ((tid >> 7) * 64 < K) {
        const int c = tid & 127, h = tid >> 7;
        const float* fr = filt + (long long)c * K;
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = h * 64 + ch * 8 + 2 * e;
                const float a = k0 < K ? fr[k0] : 0.f, b = k0 + 1 < K ? fr[k0 + 1] : 0.f;
                const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
                hi[e] = pack_bf16x2(ah, bh);
                lo[e] = pack_bf16x2(a - ah, b - bh);
            }
            const uint32_t o = stc_chunk_off(c, ch);
            *reinterpret_cast<uint4*>(smem + STC_SM_B + h * 16384 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(smem + STC_SM_B + 32768 + h * 16384 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(128, 128, false, false);

    // operand of tile `tile` into buffer `buf`: stage the 255 samples, then thread (row tid & 127, tap half tid >> 7)
    // writes its 64 taps as hi / lo bf16
    auto build = [&](int tile, int buf) {
        const int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * 128;
        const float* xb = x + (long long)b * N;
        float* sx = reinterpret_cast<float*>(smem + STC_SM_X + buf * 1024);
        if (tid < 255) {
            const int j = t0 + tid - half;
            sx[tid] = (j < N + half) ? xb[reflect_index(j, N)] : 0.f;
        }
        __syncthreads();
        const int r = tid & 127, h = tid >> 7;
        uint8_t* ab = smem + STC_SM_A + buf * STC_A_BYTES;
#pragma unroll 2
        for (int ch = 0; ch < (h * 64 < K ? 8 : 0); ++ch) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = h * 64 + ch * 8 + 2 * e;
                const float a = sx[r + k0], c = k0 + 1 < 128 ? sx[(r + k0 + 1) < 255 ? r + k0 + 1 : 254] : 0.f;
                const float ah = __bfloat162float(__float2bfloat16_rn(a)), ch_ = __bfloat162float(__float2bfloat16_rn(c));
                hi[e] = pack_bf16x2(ah, ch_);
                lo[e] = pack_bf16x2(a - ah, c - ch_);
            }
            const uint32_t o = stc_chunk_off(r, ch);
            *reinterpret_cast<uint4*>(ab + h * 16384 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(ab + 32768 + h * 16384 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();  // generic-proxy writes -> the tensor pipe's async-proxy reads
    };
    auto issue = [&](int buf) {  // thread 0, after the __syncthreads that follows build(): 3 x 8 products of 128 x 128 x 16
        const uint32_t a0 = smem_u32(smem + STC_SM_A + buf * STC_A_BYTES), b0 = smem_u32(smem + STC_SM_B);
        const uint32_t td = tmem_base + buf * 128;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            const uint32_t aa = a0 + (pr == 2 ? 32768u : 0u), bb = b0 + (pr == 1 ? 32768u : 0u);  // hi*hi, hi*lo, lo*hi
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (s >= nsteps) break;
                const uint32_t ko = (uint32_t)(s >> 2) * 16384u + (uint32_t)(s & 3) * 32u;
                umma_bf16(td, umma_smem_desc(aa + ko, 0, 1024), umma_smem_desc(bb + ko, 0, 1024), idesc,
                          (pr > 0 || s > 0) ? 1u : 0u);
            }
        }
        umma_commit(&bar_mma[buf]);
    };

    int it = 0;
    const int first = blockIdx.x;
    if (first < num_tiles) {
        build(first, 0);
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue(0);
        }
    }
    for (int tile = first; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int next = tile + gridDim.x;
        if (next < num_tiles) {
            // the other accumulator buffer was drained in the previous iteration (the __syncthreads inside build orders it)
            build(next, buf ^ 1);
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue(buf ^ 1);
            }
        }
        mbar_wait(&bar_mma[buf], (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        // epilogue: thread = output row (TMEM lane), 64 of the 128 channels per warp group
        const int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * 128;
        const int r = (warp & 3) * 32 + lane, t = t0 + r;
        const int cbase = (warp >> 2) * 64;
        bf16* yrow = y + ((long long)b * N + (t < N ? t : 0)) * SINC_CPAD + cbase;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 128 + cbase + c * 32, raw);
            tmem_ld_wait();
            if (t < N) {
#pragma unroll
                for (int d = 0; d < 32; d += 8) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(raw[d]), __uint_as_float(raw[d + 1]));
                    v.y = pack_bf16x2(__uint_as_float(raw[d + 2]), __uint_as_float(raw[d + 3]));
                    v.z = pack_bf16x2(__uint_as_float(raw[d + 4]), __uint_as_float(raw[d + 5]));
                    v.w = pack_bf16x2(__uint_as_float(raw[d + 6]), __uint_as_float(raw[d + 7]));
                    *reinterpret_cast<uint4*>(yrow + c * 32 + d) = v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// Filter gradient on tcgen05 (bf16 dy):  dfilt[c][k] += sum_{b,t} dy[b, t, c] * x[b, reflect(t + k - K/2)]
// = dy^T (channels x time) times the Toeplitz operand (time x taps) of sinc_conv_fwd_tc_kernel. Per 128-sample tile
// both operands sit in shared memory with the time index along the rows (MN-major for the MMA: M = channels, N = taps,
// K = time): dy as stored (bf16), x split into hi + lo -- two products of 8 K steps, accumulated over ALL tiles of the
// CTA in one 128 x 128 fp32 TMEM accumulator; one atomic add per (channel, tap) and CTA at the end.
// ------------------------------------------------------------------------------------------
constexpr int SWG_SM_DY = 0;                       // two buffers of [2 channel halves][128 t][64 c] bf16
constexpr int SWG_SM_A = 2 * 32768;                // two buffers of hi / lo x [2 tap halves][128 t][64 k]
constexpr int SWG_SM_X = SWG_SM_A + 2 * STC_A_BYTES;
constexpr int SWG_SM_BAR = SWG_SM_X + 2 * 1024;
constexpr int SWG_SMEM_TOTAL = SWG_SM_BAR + 64 + 1024;

__global__ void __launch_bounds__(256, 1) sinc_conv_wgrad_tc_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                                    float* __restrict__ dfilt, int N, int K,
                                                                    int tiles_per_clip, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + SWG_SM_BAR);  // [2]: the products reading buffer i retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = K / 2;
    if (tid == 0) {
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nhalf = K > 64 ? 2 : 1;  // 64-tap halves that hold real taps (the N extent of the product)
    const uint32_t idesc_n = umma_idesc_bf16(128, nhalf * 64, true, true);

    auto build = [&](int tile, int buf) {
        const int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * 128;
        const float* xb = x + (long long)b * N;
        float* sx = reinterpret_cast<float*>(smem + SWG_SM_X + buf * 1024);
        if (tid < 255) {
            const int j = t0 + tid - half;
            sx[tid] = (j < N + half) ? xb[reflect_index(j, N)] : 0.f;
        }
        // dy tile: row t (256 B = 16 chunks of 8 channels) -> [channel half][t][64 c], 128-byte swizzle; rows beyond N: zeros
        {
            uint8_t* db = smem + SWG_SM_DY + buf * 32768;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = tid + 256 * i;  // 128 rows x 16 chunks
                const int r = idx >> 4, ch = idx & 15;
                const int t = t0 + r;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (t < N) v = *reinterpret_cast<const uint4*>(dy + ((long long)b * N + t) * SINC_CPAD + ch * 8);
                *reinterpret_cast<uint4*>(db + (ch >> 3) * 16384 + stc_chunk_off(r, ch & 7)) = v;
            }
        }
        __syncthreads();
        const int r = tid & 127, h = tid >> 7;
        uint8_t* ab = smem + SWG_SM_A + buf * STC_A_BYTES;
#pragma unroll 2
        for (int ch = 0; ch < (h * 64 < K ? 8 : 0); ++ch) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = h * 64 + ch * 8 + 2 * e;
                const float a = sx[r + k0], c = sx[r + k0 + 1 < 255 ? r + k0 + 1 : 254];
                const float ah = __bfloat162float(__float2bfloat16_rn(a)), ch_ = __bfloat162float(__float2bfloat16_rn(c));
                hi[e] = pack_bf16x2(ah, ch_);
                lo[e] = pack_bf16x2(a - ah, c - ch_);
            }
            const uint32_t o = stc_chunk_off(r, ch);
            *reinterpret_cast<uint4*>(ab + h * 16384 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(ab + 32768 + h * 16384 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
    };
    auto issue = [&](int buf, bool first) {  // dy^T x_hi + dy^T x_lo: 2 x 8 K steps of 16 time rows
        const uint32_t a0 = smem_u32(smem + SWG_SM_DY + buf * 32768), b0 = smem_u32(smem + SWG_SM_A + buf * STC_A_BYTES);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int s = 0; s < 8; ++s)
                umma_bf16(tmem_base, umma_smem_desc(a0 + s * 2048, 16384, 1024),
                          umma_smem_desc(b0 + pr * 32768 + s * 2048, 16384, 1024), idesc_n, (first && pr == 0 && s == 0) ? 0u : 1u);
        umma_commit(&bar_mma[buf]);
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) {  // the products of two tiles ago read this buffer
            mbar_wait(&bar_mma[buf], (uint32_t)(((it >> 1) - 1) & 1));
        }
        build(tile, buf);
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue(buf, it == 0);
        }
    }
    if (it > 0) {
        // every product retired (commits arrive in order: the last one covers all)
        const int last = (it - 1) & 1;
        mbar_wait(&bar_mma[last], (uint32_t)((((it - 1) >> 1)) & 1));
        tc_fence_after();
        // epilogue: thread = channel (TMEM lane), warps 0-3 taps 0..63, warps 4-7 taps 64..127
        const int c = (warp & 3) * 32 + lane;
        const int kbase = (warp >> 2) * 64;
        if (kbase < nhalf * 64) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kbase + q * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int k = kbase + q * 32 + i;
                    if (k < K) atomicAdd(dfilt + (long long)c * K + k, __uint_as_float(raw[i]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<128>(tmem_base);
    }
}

// dfilt[c][k] += sum_{b,t} dy[b, t, c] * x[b, reflect(t + k - K/2)]
// grid (chunks, B); each block walks its chunk of time steps in tiles of SINC_TT.
template <typename TI>
__global__ void __launch_bounds__(256) sinc_conv_wgrad_kernel(const float* __restrict__ x, const TI* __restrict__ dy,
                                                              float* __restrict__ dfilt, int N, int K,
                                                              int steps_per_block) {
    extern __shared__ float sm[];
    float* sdy = sm;                         // [SINC_TT][128]
    float* sx = sm + SINC_TT * SINC_CPAD;    // [SINC_TT + K]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = K / 2;
    const int kpw = (K + 7) / 8;             // taps per warp (<= 16)
    const int k_begin = warp * kpw;
    float acc[16][4];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[k][c] = 0.f;
    const float* xb = x + (long long)b * N;
    const int c_begin = blockIdx.x * steps_per_block;
    const int c_end = min(N, c_begin + steps_per_block);
    for (int t0 = c_begin; t0 < c_end; t0 += SINC_TT) {
        __syncthreads();
        for (int i = threadIdx.x; i < SINC_TT * 32; i += 256) {
            const int tl = i >> 5, c4 = (i & 31) * 4;
            const int t = t0 + tl;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (t < c_end) load4(dy + ((long long)b * N + t) * SINC_CPAD + c4, v);
            *reinterpret_cast<float4*>(sdy + tl * SINC_CPAD + c4) = make_float4(v[0], v[1], v[2], v[3]);
        }
        for (int i = threadIdx.x; i < SINC_TT + K; i += 256) {
            const int j = t0 + i - half;
            sx[i] = (j < N + half) ? xb[reflect_index(j, N)] : 0.f;
        }
        __syncthreads();
        if (kpw <= 8) {
            // sliding window: the 8 taps of time step tl + 1 reuse 7 of the 8 samples of step tl, so eight time steps
            // cost 15 sample loads + 8 gradient loads for 256 FMAs (taps beyond K accumulate into unused slots)
            for (int tl = 0; tl < SINC_TT; tl += 8) {
                float xw[15];
#pragma unroll
                for (int i = 0; i < 15; ++i) xw[i] = sx[min(tl + k_begin + i, SINC_TT + K - 1)];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 d = *reinterpret_cast<const float4*>(sdy + (tl + u) * SINC_CPAD + lane * 4);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        acc[k][0] = fmaf(d.x, xw[u + k], acc[k][0]);
                        acc[k][1] = fmaf(d.y, xw[u + k], acc[k][1]);
                        acc[k][2] = fmaf(d.z, xw[u + k], acc[k][2]);
                        acc[k][3] = fmaf(d.w, xw[u + k], acc[k][3]);
                    }
                }
            }
        } else {
            for (int tl = 0; tl < SINC_TT; ++tl) {
                const float4 d = *reinterpret_cast<const float4*>(sdy + tl * SINC_CPAD + lane * 4);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k < kpw && k_begin + k < K) {
                        const float xv = sx[tl + k_begin + k];
                        acc[k][0] += d.x * xv;
                        acc[k][1] += d.y * xv;
                        acc[k][2] += d.z * xv;
                        acc[k][3] += d.w * xv;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < kpw && k_begin + k < K) {
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(dfilt + (long long)(lane * 4 + c) * K + k_begin + k, acc[k][c]);
        }
    }
}

}  // namespace a2v

using namespace a2v;

static int check_sinc_args(int C, int K) {
    A2V_REQUIRE(C > 0 && C <= SINC_CPAD, "sinc: out_channels must be in (0, 128], got %d", C);
    A2V_REQUIRE(K >= 3 && (K & 1) == 1 && K <= SINC_KMAX, "sinc: kernel_size must be odd and <= %d, got %d", SINC_KMAX, K);
    return A2V_OK;
}

extern "C" int a2v_sinc_filters_fwd(const float* low_hz, const float* band_hz, const float* n_, const float* window_,
                                    int C, int K, float min_low_hz, float min_band_hz, float sample_rate,
                                    float* filters, a2v_stream_t stream) {
    int rc = check_sinc_args(C, K);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(low_hz && band_hz && n_ && window_ && filters, "sinc_filters_fwd: NULL pointer");
    sinc_filters_fwd_kernel<<<SINC_CPAD, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        low_hz, band_hz, n_, window_, C, K, min_low_hz, min_band_hz, sample_rate * 0.5f, filters);
    return a2v_check_launch("sinc_filters_fwd");
}

extern "C" int a2v_sinc_filters_bwd(const float* low_hz, const float* band_hz, const float* n_, const float* window_,
                                    int C, int K, float min_low_hz, float min_band_hz, float sample_rate,
                                    const float* dfilters, float* dlow_hz, float* dband_hz, a2v_stream_t stream) {
    int rc = check_sinc_args(C, K);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(low_hz && band_hz && n_ && window_ && dfilters && dlow_hz && dband_hz, "sinc_filters_bwd: NULL pointer");
    sinc_filters_bwd_kernel<<<C, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        low_hz, band_hz, n_, window_, C, K, min_low_hz, min_band_hz, sample_rate * 0.5f, dfilters, dlow_hz, dband_hz);
    return a2v_check_launch("sinc_filters_bwd");
}

extern "C" int a2v_sinc_conv_fwd(int out_dtype, const float* x, const float* filters, void* y, int B, int N, int K,
                                 a2v_stream_t stream) {
    int rc = check_sinc_args(1, K);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(x && filters && y && B > 0 && N > K, "sinc_conv_fwd: bad arguments");
    A2V_REQUIRE(out_dtype == A2V_F32 || out_dtype == A2V_BF16, "sinc_conv_fwd: bad dtype");
    const size_t smem = (size_t)(K * SINC_CPAD + SINC_TT + K + 3) * sizeof(float);
    dim3 grid(ceil_div(N, SINC_TT), B);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_dtype == A2V_F32) {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_fwd_kernel<float>), 100 * 1024) != A2V_OK) return A2V_ERR_CUDA;
        sinc_conv_fwd_kernel<float><<<grid, 256, smem, st>>>(x, filters, (float*)y, N, K);
    } else {
        static int use_tc = -1;  // A2V_SINC_TC=0: CUDA-core kernel for the bf16 output as well
        if (use_tc < 0) {
            const char* e = getenv("A2V_SINC_TC");
            use_tc = (e == nullptr || e[0] != '0') ? 1 : 0;
        }
        if (use_tc == 1 && K <= 128 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
            if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_fwd_tc_kernel), STC_SMEM_TOTAL) != A2V_OK)
                return A2V_ERR_CUDA;
            const int tiles_per_clip = ceil_div(N, 128), num_tiles = tiles_per_clip * B;
            const int g = num_tiles < a2v_num_sms() ? num_tiles : a2v_num_sms();
            sinc_conv_fwd_tc_kernel<<<g, 256, STC_SMEM_TOTAL, st>>>(x, filters, (bf16*)y, N, K, tiles_per_clip, num_tiles);
            return a2v_check_launch("sinc_conv_fwd_tc");
        }
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_fwd_kernel<bf16>), 100 * 1024) != A2V_OK) return A2V_ERR_CUDA;
        sinc_conv_fwd_kernel<bf16><<<grid, 256, smem, st>>>(x, filters, (bf16*)y, N, K);
    }
    return a2v_check_launch("sinc_conv_fwd");
}

extern "C" int a2v_sinc_conv_wgrad(int dy_dtype, const float* x, const void* dy, float* dfilters, int B, int N, int K,
                                   a2v_stream_t stream) {
    int rc = check_sinc_args(1, K);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(x && dy && dfilters && B > 0 && N > K, "sinc_conv_wgrad: bad arguments");
    A2V_REQUIRE(dy_dtype == A2V_F32 || dy_dtype == A2V_BF16, "sinc_conv_wgrad: bad dtype");
    const size_t smem = (size_t)(SINC_TT * SINC_CPAD + SINC_TT + K) * sizeof(float);
    int chunks = (a2v_num_sms() * 2) / B;
    if (chunks < 1) chunks = 1;
    int steps = ceil_div(ceil_div(N, chunks), SINC_TT) * SINC_TT;
    dim3 grid(ceil_div(N, steps), B);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dy_dtype == A2V_F32) {
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_wgrad_kernel<float>), 100 * 1024) != A2V_OK) return A2V_ERR_CUDA;
        sinc_conv_wgrad_kernel<float><<<grid, 256, smem, st>>>(x, (const float*)dy, dfilters, N, K, steps);
    } else {
        static int use_tc = -1;
        if (use_tc < 0) {
            const char* e = getenv("A2V_SINC_TC");
            use_tc = (e == nullptr || e[0] != '0') ? 1 : 0;
        }
        if (use_tc == 1 && K <= 128 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
            if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_wgrad_tc_kernel), SWG_SMEM_TOTAL) != A2V_OK)
                return A2V_ERR_CUDA;
            const int tiles_per_clip = ceil_div(N, 128), num_tiles = tiles_per_clip * B;
            const int g = num_tiles < a2v_num_sms() ? num_tiles : a2v_num_sms();
            sinc_conv_wgrad_tc_kernel<<<g, 256, SWG_SMEM_TOTAL, st>>>(x, (const bf16*)dy, dfilters, N, K, tiles_per_clip, num_tiles);
            return a2v_check_launch("sinc_conv_wgrad_tc");
        }
        if (a2v_ensure_dynamic_smem(reinterpret_cast<const void*>(sinc_conv_wgrad_kernel<bf16>), 100 * 1024) != A2V_OK) return A2V_ERR_CUDA;
        sinc_conv_wgrad_kernel<bf16><<<grid, 256, smem, st>>>(x, (const bf16*)dy, dfilters, N, K, steps);
    }
    return a2v_check_launch("sinc_conv_wgrad");
}
