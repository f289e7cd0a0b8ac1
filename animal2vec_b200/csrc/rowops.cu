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

template <typename T, int NCH>
__global__ void __launch_bounds__(256) rowln_fwd_kernel(const RowLnParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int C = p.C;
    const bool padded = p.gr < p.gw;
    int creal = (C / p.gw) * p.gr;
    const float inv_c = 1.0f / (float)creal;
    const float keep_b = p.drop_b > 0.f ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    const float keep_o = p.drop_out > 0.f ? 1.0f / (1.0f - p.drop_out) : 1.0f;
    const T* A = reinterpret_cast<const T*>(p.a);
    const T* Bp = reinterpret_cast<const T*>(p.b);
    const T* P = reinterpret_cast<const T*>(p.post);
    T* Y = reinterpret_cast<T*>(p.y);

    for (long long row = warp0; row < p.rows; row += nwarps) {
        float z[NCH][4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                load4(A + row * C + c, z[i]);
                if (Bp != nullptr) {
                    float t[4];
                    load4(Bp + row * C + c, t);
                    if (p.drop_b > 0.f) {
                        bool k[4];
                        drop_keep4(p.seed_b, (unsigned long long)(row * C + c) >> 2, p.drop_b, k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = k[j] ? t[j] * keep_b : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[i][j] += t[j];
                }
                if (padded) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (((c + j) % p.gw) >= p.gr) z[i][j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) s += z[i][j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) z[i][j] = 0.f;
            }
        }
        const float mean = warp_sum(s) * inv_c;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = !padded || ((c + j) % p.gw) < p.gr;
                    const float d = z[i][j] - mean;
                    v += real ? d * d : 0.f;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(v) * inv_c + p.eps);
        if (lane == 0) {
            if (p.mean != nullptr) p.mean[row] = mean;
            if (p.rstd != nullptr) p.rstd[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                float o[4];
                bool ko[4] = {true, true, true, true};
                if (p.drop_out > 0.f) drop_keep4(p.seed_out, (unsigned long long)(row * C + c) >> 2, p.drop_out, ko);
                float post[4] = {0.f, 0.f, 0.f, 0.f};
                if (P != nullptr) load4(P + row * C + c, post);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = !padded || ((c + j) % p.gw) < p.gr;
                    const int pi = padded ? param_index(c + j, p.gw, p.gr) : (c + j);
                    float n = (z[i][j] - mean) * rstd;
                    if (p.gamma != nullptr) n = n * p.gamma[real ? pi : 0] + (p.beta != nullptr ? p.beta[real ? pi : 0] : 0.f);
                    float y;
                    if (p.act == 1) {
                        y = gelu_exact(n);
                    } else if (p.act == 2) {
                        y = n * p.act_alpha[real ? pi : 0] * sigmoidf_(p.act_beta[real ? pi : 0] * n);
                    } else {
                        y = n;
                    }
                    if (p.drop_out > 0.f) y = ko[j] ? y * keep_o : 0.f;
                    y += post[j];
                    o[j] = real ? y : 0.f;
                }
                store4(Y + row * C + c, o);
            }
        }
    }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256) rowln_bwd_kernel(const RowLnParams p) {
    extern __shared__ float sred[];  // [4][C] partial parameter gradients
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int C = p.C;
    const bool padded = p.gr < p.gw;
    const int creal = (C / p.gw) * p.gr;
    const float inv_c = 1.0f / (float)creal;
    const float keep_b = p.drop_b > 0.f ? 1.0f / (1.0f - p.drop_b) : 1.0f;
    const float keep_o = p.drop_out > 0.f ? 1.0f / (1.0f - p.drop_out) : 1.0f;
    const T* A = reinterpret_cast<const T*>(p.a);
    const T* Bp = reinterpret_cast<const T*>(p.b);
    const T* DY = reinterpret_cast<const T*>(p.dy);
    T* DA = reinterpret_cast<T*>(p.da);
    T* DB = reinterpret_cast<T*>(p.db);
    const bool want_affine = p.dgamma != nullptr;
    const bool want_act = p.dact_alpha != nullptr;

    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();

    float acc_g[NCH][4], acc_b[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_g[i][j] = acc_b[i][j] = 0.f;

    for (long long row = warp0; row < p.rows; row += nwarps) {
        float xh[NCH][4], g[NCH][4];
        const float mean = p.mean[row], rstd = p.rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) xh[i][j] = g[i][j] = 0.f;
            if (c < C) {
                float z[4], dy[4];
                load4(A + row * C + c, z);
                if (Bp != nullptr) {
                    float t[4];
                    load4(Bp + row * C + c, t);
                    if (p.drop_b > 0.f) {
                        bool k[4];
                        drop_keep4(p.seed_b, (unsigned long long)(row * C + c) >> 2, p.drop_b, k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = k[j] ? t[j] * keep_b : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[j] += t[j];
                }
                load4(DY + row * C + c, dy);
                if (p.drop_out > 0.f) {
                    bool ko[4];
                    drop_keep4(p.seed_out, (unsigned long long)(row * C + c) >> 2, p.drop_out, ko);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dy[j] = ko[j] ? dy[j] * keep_o : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = !padded || ((c + j) % p.gw) < p.gr;
                    if (!real) continue;
                    const int pi = padded ? param_index(c + j, p.gw, p.gr) : (c + j);
                    const float x = (z[j] - mean) * rstd;
                    const float gam = p.gamma != nullptr ? p.gamma[pi] : 1.f;
                    const float n = x * gam + (p.beta != nullptr ? p.beta[pi] : 0.f);
                    float dn;
                    if (p.act == 1) {
                        dn = dy[j] * gelu_exact_grad(n);
                    } else if (p.act == 2) {
                        const float al = p.act_alpha[pi], be = p.act_beta[pi];
                        const float sg = sigmoidf_(be * n);
                        dn = dy[j] * al * (sg + n * be * sg * (1.f - sg));
                        if (want_act) {
                            // reuse acc_* of the (unused when act==2 has its own) slots below
                            atomicAdd(&sred[2 * C + c + j], dy[j] * n * sg);
                            atomicAdd(&sred[3 * C + c + j], dy[j] * al * n * n * sg * (1.f - sg));
                        }
                    } else {
                        dn = dy[j];
                    }
                    if (want_affine) {
                        acc_g[i][j] += dn * x;
                        acc_b[i][j] += dn;
                    }
                    xh[i][j] = x;
                    g[i][j] = dn * gam;
                    s1 += g[i][j];
                    s2 += g[i][j] * x;
                }
            }
        }
        s1 = warp_sum(s1) * inv_c;
        s2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                float dz[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool real = !padded || ((c + j) % p.gw) < p.gr;
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
            if (c < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&sred[c + j], acc_g[i][j]);
                    atomicAdd(&sred[C + c + j], acc_b[i][j]);
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

template <typename T>
static int launch_rowln(const RowLnParams& p, bool bwd, cudaStream_t st) {
    const int nch = ceil_div(p.C, 128);
    const int threads = 256;
    long long blocks_needed = ceil_div64(p.rows, threads / 32);
    int grid = (int)(blocks_needed < (long long)a2v_num_sms() * 8 ? blocks_needed : (long long)a2v_num_sms() * 8);
    if (grid < 1) grid = 1;
    const size_t smem = bwd ? (size_t)4 * p.C * sizeof(float) : 0;
#define A2V_ROWLN(N)                                                          \
    do {                                                                      \
        if (bwd)                                                              \
            rowln_bwd_kernel<T, N><<<grid, threads, smem, st>>>(p);          \
        else                                                                  \
            rowln_fwd_kernel<T, N><<<grid, threads, 0, st>>>(p);             \
    } while (0)
    if (nch <= 1) A2V_ROWLN(1);
    else if (nch <= 2) A2V_ROWLN(2);
    else if (nch <= 4) A2V_ROWLN(4);
    else if (nch <= 6) A2V_ROWLN(6);
    else A2V_ROWLN(8);
#undef A2V_ROWLN
    return a2v_check_launch(bwd ? "rowln_bwd" : "rowln_fwd");
}

static int validate_rowln(const a2v_rowln_desc* d) {
    A2V_REQUIRE(d != nullptr, "rowln: NULL descriptor");
    A2V_REQUIRE(d->rows >= 0 && d->channels > 0 && d->channels % 4 == 0 && d->channels <= 1024,
                "rowln: channels must be a multiple of 4 in (0, 1024], got %d", d->channels);
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
    return d->dtype == A2V_F32 ? launch_rowln<float>(p, false, st) : launch_rowln<bf16>(p, false, st);
}

extern "C" int a2v_rowln_bwd(const a2v_rowln_desc* d, a2v_stream_t stream) {
    int rc = validate_rowln(d);
    if (rc != A2V_OK) return rc;
    A2V_REQUIRE(d->dy != nullptr && d->mean != nullptr && d->rstd != nullptr,
                "rowln_bwd: dy / saved mean / saved rstd are required");
    A2V_REQUIRE(d->db == nullptr || d->b != nullptr, "rowln_bwd: db requested without b");
    if (d->rows == 0) return A2V_OK;
    RowLnParams p = to_params(d);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return d->dtype == A2V_F32 ? launch_rowln<float>(p, true, st) : launch_rowln<bf16>(p, true, st);
}
