/*
 * animal2vec_b200 C-ABI  --  liba2v_sm100.so
 *
 * The reference (livingingroups/animal2vec) is pure Python/PyTorch: it has no native
 * FFI of its own. Each entry point below replaces the *PyTorch library dispatch* at the
 * cited reference call site (file:line under /root/reference) with a hand-written
 * sm_100a kernel. SURVEY.md section 8(b) fixes the conventions:
 *   - plain pointers and sizes only, no torch types;
 *   - the caller owns every buffer (device memory, workspaces); the library never
 *     allocates device memory, never synchronises, never changes the current device;
 *   - every call takes the CUDA stream it must launch on;
 *   - return 0 on success, non-zero a2v status otherwise; a2v_last_error() gives the
 *     thread-local message. Shape/alignment violations are rejected before launch.
 *
 * dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef A2V_CAPI_H
#define A2V_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* a2v_stream_t; /* cudaStream_t */

#define A2V_F32 0
#define A2V_BF16 1

const char* a2v_last_error(void);
int a2v_version(void);
/* 1 iff the current device is compute capability 10.x (sm_100 family). */
int a2v_device_supported(void);
int a2v_num_sms(void);

/* ------------------------------------------------------------------------------------
 * tcgen05 / TMA / TMEM GEMM  (replaces every nn.Linear / nn.Conv1d GEMM on the path:
 * nn/modalities/modules.py:356,371-374,408 (qkv/proj), timm Mlp fc1/fc2 via
 * modules.py:312-317, nn/modalities/audio.py:87 (project_features), audio.py:97-103
 * (grouped positional conv), modules.py:143-149 (decoder grouped conv), modules.py:173
 * (decoder proj), nn/utils.py:1085-1090 (feature-extractor convs) and their autograd
 * backward GEMMs).
 *
 * C(m, n) = alpha * sum_k A(m, k) * B(n, k)   [+ epilogue], bf16 operands, fp32 accumulate.
 *
 * mode 0 ("NT", both operands K-major): A is a (K-inner, rows, batch) view, B is a
 *   (K-inner, N-rows) matrix. The K loop runs over `taps` x `k_per_tap`; tap j reads A rows
 *   shifted by a_row_off + j * a_tap_rows (out-of-range rows read as zeros through TMA
 *   bounds checking): that is a stride-1 (grouped) 1-D convolution without im2col. groups > 1
 *   gives a block-diagonal product (group g uses A columns g*a_group_stride.., B rows
 *   g*b_group_stride.., C columns g*c_group_stride..).
 * mode 1 ("TN", both operands MN-major): C(m, n) = sum over (batch, row) of
 *   A[batch, row, m] * B[batch, row + tap*b_tap_rows + b_row_off, n]: weight gradients.
 *   The reduction may be split `k_splits` ways; partial sums are combined with fp32 atomics
 *   (out_atomic = 1, C must then be fp32 and pre-initialised, which also gives gradient
 *   accumulation for free).
 * ------------------------------------------------------------------------------------ */
typedef struct {
    const void* ptr;       /* bf16 */
    int64_t dim0;          /* innermost (contiguous) extent, elements */
    int64_t dim1;          /* rows */
    int64_t dim2;          /* batch (1 for a plain matrix) */
    int64_t stride1;       /* elements between rows (multiple of 8) */
    int64_t stride2;       /* elements between batches (multiple of 8) */
} a2v_operand;

typedef struct {
    int mode;              /* 0 NT, 1 TN */
    int block_n;           /* 64, 128 or 256 */
    a2v_operand a, b;
    int M, N;              /* output block extent per (batch, group[, tap]) */
    int k_per_tap;         /* NT: K elements per tap (multiple of 8; TMA zero-fills the tail) */
    int taps;
    int batch;
    int groups;
    int a_group_stride, a_row_off, a_tap_rows;
    int b_group_stride, b_row_off, b_tap_rows;
    int red_rows;          /* TN: rows per batch that are reduced */
    int k_splits;          /* TN: reduction splits (>=1) */
    void* c;
    int c_dtype;           /* A2V_F32 / A2V_BF16 */
    int out_atomic;        /* 1: atomically add into fp32 C */
    int out_accumulate;    /* 1: C += result (non-atomic, fp32 only) */
    int64_t ldc;
    int64_t c_batch_stride;/* rows */
    int64_t c_row_off;
    int c_group_stride;    /* NT: columns per group; TN: rows per group */
    int c_tap_stride;      /* TN: columns per tap */
    float alpha;
    const float* bias;     /* NT: per output column (global column index), or NULL */
    int act;               /* 0 none, 1 exact GELU */
    void* preact;          /* optional: pre-activation copy, same layout/dtype as C */
    const void* residual;  /* optional: added after the activation, same layout/dtype as C */
    const void* dgelu_u;   /* optional: result *= GELU'(u), same layout/dtype as C */
    int a_tap_cols;        /* TN only, 0 = off: M index = tap * a_tap_cols + channel; the A box of tap j reads
                              rows shifted by a_row_off + j * a_tap_rows (conv weight gradient with x as the
                              M side: C[(j, c), n] = sum_t x[t + j - pad, c] * dy[t, n]) */
    /* STRIDED Conv1d without im2col / col2im (nn/utils.py:1085-1092: Conv1d(k, stride s, padding ceil(s/2))): the
     * channels-last input (B, T, C) is viewed as (B, T/s, s*C), so input row s*t + q is (row t + floor(q/s),
     * column block q mod s). a_tap_wrap = s > 0 (with a_tap_cols = C) turns that on for the A operand:
     *   NT (forward): tap j reads row m + floor((j + a_row_off)/s), columns ((j + a_row_off) mod s) * C + k, a_row_off = -pad;
     *   TN (weight gradient, x as the M side): the A box of tap j likewise.
     * Data gradient = grouped NT product over the s output column blocks r (dx viewed (B, T/s, s*C)); group r reads
     * dy rows m + (r + a_grow_add) / a_grow_div - u for its tap u (a_tap_rows = -1): a_grow_div > 0 adds that per-group
     * row shift (a_grow_add = pad, a_grow_div = s). */
    int a_tap_wrap;
    int a_grow_add;
    int a_grow_div;
    /* column distance between consecutive column blocks of the wrapped tap addressing when it differs from a_tap_cols
     * (0 = a_tap_cols). Gathered convolution: rows pre-gathered as (M, taps, C) -- A viewed (M, taps*C), a_tap_wrap = taps,
     * a_row_off = 0, a_tap_col_stride = C, groups select 64-channel slices inside each tap block (the last positional-conv
     * layer evaluated only on the rows the student keeps, nn/modalities/base.py:278-280). */
    int a_tap_col_stride;
    /* optional, with dgelu_u on a plain bf16 Linear product: += column sums of the stored result, fp32 (N) -- the bias
     * gradient of the Linear in front of the activation (timm Mlp.fc1, nn/modalities/modules.py:312-317), so the
     * activation backward needs no pass of its own */
    float* colsum;
} a2v_gemm_desc;

int a2v_gemm(const a2v_gemm_desc* d, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Grouped stride-1 tap convolution with 64-channel groups, activation rows loaded once per tile
 * ("slab" implicit GEMM, tcgen05/TMA): forward and data gradient of the positional-encoder convs
 * (nn/modalities/audio.py:93-113) and of Decoder1d's convs (nn/modalities/modules.py:141-157).
 *   y[b, t, g*y_group_cols + n] = sum_{j<taps, c<64} x[b, t+j-pad, g*64+c] * w[g*w_group_rows+n, j*64+c] (+bias)
 * x: (batch, T, ldx) bf16, w: (groups*w_group_rows, ldw) bf16 tap-major rows, y: (batch, T, ldy) bf16/fp32.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    const void* x;
    const void* w;
    void* y;
    int batch, T, groups, taps, pad;
    int ng;               /* outputs per group, <= 64 */
    int x_group_cols;     /* channel distance between the groups of x: 64, or a smaller multiple of 16 with x_real_cols <=
                             x_group_cols (compact group-padded layout; the 64-channel box then overlaps the next group) */
    int w_group_rows;
    int y_group_cols;
    int64_t ldx, ldw, ldy;
    int y_dtype;
    const float* bias;    /* per global y column, or NULL */
    int x_real_cols;      /* conv_slab_fwd: channels of each 64-wide input group that can be non-zero (0 = all 64);
                             the K steps over the rest are skipped (group-padded decoder layout: 48 of 64) */
} a2v_conv_desc;

int a2v_conv_slab_supported(const a2v_conv_desc* d);
int a2v_conv_slab_fwd(const a2v_conv_desc* d, a2v_stream_t stream);
/* weight gradient of the same operator: d->x = activations, d->w = dy (batch, T, ldw) with w_group_rows =
 * dy columns per group; out[(g*taps + j)*64 + c, n] += sum_{b,t} x[b, t+j-pad, g*64+c] * dy[b, t, g*w_group_rows+n]
 * (fp32, (groups*taps*64, ldo), atomically accumulated: split-K over (batch, time)). */
int a2v_conv_slab_wgrad(const a2v_conv_desc* d, float* out, int64_t ldo, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fused row LayerNorm family (HBM-bound, warp-per-row):
 *   z = a + dropout_b(b);  n = LN(z) * gamma + beta;  y = dropout_out(act(n)) + post
 * Replaces: Fp32LayerNorm+PSwish / +GELU of the feature extractor (nn/utils.py:1105-1117,
 * 1413-1435), project_features' Fp32LayerNorm (nn/modalities/audio.py:86), LayerNorm+GELU
 * of the positional-conv and decoder stacks (audio.py:104-108, nn/modalities/modules.py:
 * 150-157, residual 124-134), AltBlock's post-LN residual norms (modules.py:329-333) and
 * BlockEncoder's LN -> dropout (modules.py:84-87).
 * channels-last rows of `channels` stored elements; a "group-padded" layout has
 * `group_real` meaningful channels in every `group_width` stored ones (pads are kept 0;
 * gamma/beta/act parameters are indexed by the real channel index).
 * act: 0 none, 1 exact GELU, 2 PSwish (n * alpha_c * sigmoid(beta_c * n)).
 * Dropout masks come from a counter-based hash of (seed, element index) and are
 * regenerated by the backward pass, never stored.
 * Backward: da = dL/dz, db = dropout_b-masked da, parameter gradients are atomically
 * accumulated (fp32) into dgamma/dbeta/dact_alpha/dact_beta when non-NULL. `mean`/`rstd`
 * (fp32 per row) are written by the forward and read by the backward.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    int dtype;                 /* of a, b, post, y, dy, da, db */
    int64_t rows;
    int channels, group_width, group_real;
    float eps;
    int act;
    const void* a;
    const void* b;             /* optional */
    const float* gamma;        /* optional */
    const float* beta;         /* optional */
    const float* act_alpha;    /* PSwish */
    const float* act_beta;
    const void* post;          /* optional, added after the activation */
    void* y;
    float* mean;
    float* rstd;
    float drop_b;
    uint64_t seed_b;
    float drop_out;
    uint64_t seed_out;
    /* backward only */
    const void* dy;
    void* da;
    void* db;
    float* dgamma;
    float* dbeta;
    float* dact_alpha;
    float* dact_beta;
    float* dbias_b;            /* optional: += column sums of db (the bias gradient of the Linear that produced b) */
} a2v_rowln_desc;

int a2v_rowln_fwd(const a2v_rowln_desc* d, a2v_stream_t stream);
int a2v_rowln_bwd(const a2v_rowln_desc* d, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Masking and multi-mask cloning (HBM-bound data movement).
 *  a2v_mask_index: from the uint8 mask (rows = B*clones, 1 = masked) build
 *    ids_keep [rows,Tk], ids_restore [rows,T] (MaskInfo of nn/modalities/base.py:427-455;
 *    kept positions in ascending order) and the flat int32 row maps the gathers use:
 *      clone_src[r*T+t]       = (r/clones)*T+t if kept else -1   (clone + zero mask, base.py:244,464)
 *      keep_src_x[r*Tk+k]     = (r/clones)*T+ids_keep             (x_unmasked gather from the un-cloned x)
 *      keep_src_clone[r*Tk+k] = r*T+ids_keep                      (gather_unmasked, base.py:537-542)
 *      restore_src[r*T+t]     = r*Tk+rank if kept else -1         (decoder_input scatter, base.py:178-179)
 *    err_flag is set to 1+row if a row does not keep exactly Tk positions.
 *  a2v_row_gather: dst[o,:] = (idx[o] >= 0 ? dropout(src[idx[o],:]) : N(0, fill_std) or 0) + add[o,:].
 *  a2v_clone_sum_bwd: gradient of clone+mask+gather summed over the clones of each clip.
 * ------------------------------------------------------------------------------------ */
int a2v_mask_index(const uint8_t* mask, int rows, int T, int Tk, int clones, int32_t* ids_keep,
                   int32_t* ids_restore, int32_t* clone_src, int32_t* keep_src_x, int32_t* keep_src_clone,
                   int32_t* restore_src, int32_t* err_flag, a2v_stream_t stream);
int a2v_row_gather(int dtype, const void* src, const int32_t* idx, const void* add, void* dst, int64_t n_dst,
                   int D, float fill_std, uint64_t fill_seed, float drop_p, uint64_t drop_seed, int drop_by_src,
                   a2v_stream_t stream);
int a2v_clone_sum_bwd(int dtype, const void* d_masked, const void* d_unmasked, const int32_t* restore_src,
                      void* dx, int64_t B, int T, int clones, int D, a2v_stream_t stream);
/* Neighbourhood row map of the kept tokens: out[(r * Tk + i) * taps + j] = r * T + ids_keep[r, i] + j - pad, or -1 when
 * that frame lies outside [0, T) (zero padding of the conv). Feeds a2v_row_gather to build the (rows, taps, C) operand of a
 * convolution evaluated only at the kept positions. */
int a2v_neigh_index(const int32_t* ids_keep, int rows, int Tk, int T, int taps, int pad, int32_t* out, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Teacher targets and masked regression loss (HBM-bound).
 *  a2v_target_stats / a2v_target_apply: nn/data2vec2.py:1023-1066 (instance-norm over time of
 *    the top-K teacher FFN outputs, averaged). layers_dev = device array of K pointers to
 *    (B,T,D) tensors; stats = (K,B,D,2) fp32 (mean, rstd); y = (B,T,D) fp32.
 *  a2v_d2v_loss_fwd: data2vec2.py:850-862,1005-1021 + compute_var 1095-1110. pred (R,T,D) with
 *    R = B*clones, y (B,T,D) fp32, mask (R,T) uint8. Adds scale*sum((pred-y)^2 over masked
 *    rows) into loss_sum[0] (double) and the column sums [sum x, sum x^2, sum y, sum y^2]
 *    over masked rows into colstats (4,D) doubles. Both must be zeroed by the caller.
 *  a2v_d2v_loss_bwd: dpred = mask ? 2*scale*g*(pred-y) : 0, g = *grad_out_dev (or 1 if NULL).
 * ------------------------------------------------------------------------------------ */
int a2v_target_stats(int dtype, const void* const* layers_dev, int K, int B, int T, int D, float eps, float* stats,
                     a2v_stream_t stream);
int a2v_target_apply(int dtype, const void* const* layers_dev, int K, int B, int T, int D, const float* stats,
                     float* y, a2v_stream_t stream);
int a2v_d2v_loss_fwd(int dtype, const void* pred, const float* y, const uint8_t* mask, int64_t R, int T, int clones,
                     int D, float scale, double* loss_sum, double* colstats, a2v_stream_t stream);
int a2v_d2v_loss_bwd(int dtype, const void* pred, const float* y, const uint8_t* mask, void* dpred, int64_t R, int T,
                     int clones, int D, float scale, const float* grad_out_dev, a2v_stream_t stream);
/* Forward and backward of the masked regression loss in ONE pass (bf16, D % 256 == 0): loss_sum / colstats as in
 * a2v_d2v_loss_fwd, and dpred = grad_coef * (pred - y) on masked rows, 0 elsewhere (grad_coef = 2 * scale * upstream
 * gradient). dpred may alias pred (the prediction is not needed afterwards: nn/data2vec2.py:850-862 only feeds the
 * loss) or be NULL (statistics only). The target row of a frame is read once for all `clones` clones. */
int a2v_d2v_loss_fused(int dtype, const void* pred, const float* y, const uint8_t* mask, void* dpred, int64_t R, int T,
                       int clones, int D, float scale, float grad_coef, double* loss_sum, double* colstats,
                       a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Utilities.
 *  a2v_colsum: out[c] += sum_rows x[r,c] (bias gradients).
 *  a2v_cast_strided: dense 4-D out[i0,i1,i2,i3] = in[in_offset + sum i_d*in_strides[d]] with
 *    dtype conversion (weight re-layouts: (O,I,k) -> (O,k,I), flips, transposes).
 *  a2v_split3: bf16 hi/lo split for the validation-precision GEMM (see utilops.cu).
 *  a2v_ema_step: fused EMA teacher update replacing fairseq EMAModule.step + reload
 *    (called from nn/data2vec2.py:408): shadow = d*shadow + (1-d)*student, bf16 copy out.
 *  a2v_adamw_step: fairseq Adam with decoupled weight decay; grads are multiplied by the
 *    device scalar *grad_scale (clip coefficient x 1/sample_size); writes the bf16 copy. wd_mask
 *    (one byte per 4 parameters, or NULL) switches the decay off for the weight_decay_scale-0
 *    group of nn/data2vec2.py:318-322 so that the whole flat buffer is one launch.
 *  a2v_sumsq / a2v_clip_coef: gradient-norm clipping without a host round trip.
 * ------------------------------------------------------------------------------------ */
int a2v_colsum(int dtype, const void* x, float* out, int64_t rows, int C, a2v_stream_t stream);
/* out = dh * GELU'(u) elementwise (backward of timm Mlp's GELU, nn/modalities/modules.py:312-317). */
int a2v_dgelu_mul(int dtype, const void* dh, const void* u, void* out, int64_t n, a2v_stream_t stream);
/* The same over a (rows x C) tensor with colsum[c] += sum over rows of out[:, c] fused in (fc1.bias gradient). */
int a2v_dgelu_mul_colsum(int dtype, const void* dh, const void* u, void* out, int64_t rows, int C, float* colsum,
                         a2v_stream_t stream);
int a2v_cast_strided(int in_dtype, int out_dtype, const void* in, void* out, const int64_t* dims4,
                     const int64_t* in_strides4, int64_t in_offset, a2v_stream_t stream);
/* general 4-D re-layout with cast: out[out_offset + i.out_strides] (+)= in[in_offset + i.in_strides]
 * (weight packing into tap-major / group-padded / transposed GEMM layouts and back for the
 * weight gradients; the reference keeps torch's (out, in/groups, k) Conv1d layout,
 * nn/modalities/audio.py:97-103, modules.py:143-149, nn/utils.py:1085-1090). */
int a2v_relayout(int in_dtype, int out_dtype, const void* in, void* out, const int64_t* dims4,
                 const int64_t* in_strides4, int64_t in_offset, const int64_t* out_strides4, int64_t out_offset,
                 int accumulate, a2v_stream_t stream);
/* Batched form: one launch over a DEVICE-resident table of items (same semantics per item; ``zero_src`` also
 * clears every source element after it was read, which is how the packed weight-gradient buffers are emptied
 * when they are folded into the checkpoint-layout gradient). Items must not alias each other's outputs and
 * each item must have fewer than 2^31 elements. The caller owns the table. */
typedef struct {
    const void* in;
    void* out;
    int64_t dims[4];
    int64_t in_strides[4];
    int64_t out_strides[4];
    int64_t in_offset, out_offset;
    int32_t in_dtype, out_dtype;
    int32_t accumulate, zero_src;
} a2v_relayout_item;
int a2v_relayout_batch(const a2v_relayout_item* items_device, int n_items, int blocks_per_item, a2v_stream_t stream);
int a2v_cast_f32_to_bf16(const float* in, void* out, int64_t n, a2v_stream_t stream);
/* inverse widening copy (n multiple of 8): the bf16 gradient buckets after the NCCL all-reduce (trainer.BucketReducer) */
int a2v_cast_bf16_to_f32(const void* in, float* out, int64_t n, a2v_stream_t stream);
int a2v_split3(const float* in, void* out, int64_t rows, int K, int pattern, a2v_stream_t stream);
int a2v_ema_step(const float* student, float* shadow, void* teacher_bf16, int64_t n, float decay,
                 a2v_stream_t stream);
int a2v_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* grad_scale,
                   const uint8_t* wd_mask, a2v_stream_t stream);
int a2v_sumsq(const float* x, int64_t n, double* out, a2v_stream_t stream);
int a2v_clip_coef(const double* sumsq, const float* denom, float numer, float max_norm, float* out2,
                  a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Self-attention with on-the-fly ALiBi (replaces AltAttention.forward,
 * nn/modalities/modules.py:368-410, and the whole bias-tensor machinery of
 * nn/modalities/base.py:293-314,553-698).
 *   qkv  : (batch, L, 3*H*64) rows [q | k | v], each (H, 64)   (output of the qkv Linear)
 *   out  : (batch, L, H*64);  lse : (batch, H, L) fp32 log-sum-exp (needed by the backward)
 *   pos  : (batch, L) int32 absolute token positions (ids_keep) or NULL for 0..L-1
 *   bias[h,i,j] = -slopes[h] * max(alibi_scale[h*stride], 0) * |pos_i - pos_j|
 *   dropout on the attention probabilities uses a counter-based hash of (seed, b, h, i, j).
 * bf16: tcgen05/TMA flash-style forward; backward keeps one head resident in shared memory
 * (L <= 160, the student's kept-token count). fp32: CUDA-core validation kernels; the fp32
 * backward accumulates dK/dV atomically, so dqkv must be zeroed by the caller.
 * dalibi_scale[h*stride] is atomically accumulated.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    int dtype;
    int batch, L, H, head_dim;
    const void* qkv;
    void* out;
    float* lse;
    const int32_t* pos;
    const float* slopes;
    const float* alibi_scale;
    int alibi_scale_stride;
    float sm_scale;
    float drop_p;
    uint64_t seed;
    /* backward only */
    const void* dout;
    void* dqkv;
    float* dalibi_scale;
    /* forward, optional (bf16, pos == NULL): (batch*H) pairs {max|q|^2, max|k|^2} from a2v_attn_qk_bound (which
     * accumulates with atomic max into a buffer the caller zero-fills). With it the
     * forward skips key tiles that ALiBi pushes below 2^-50 of the row maximum (their sum is far below fp32
     * resolution; the reference materialises and rounds them away, nn/modalities/modules.py:393-399). */
    const float* qk_bound;
    /* backward, bf16: 0 = automatic (L <= 160: the shared-memory-resident kernel of the short student sequences;
     * longer: the tiled kernel), 1 = resident, 2 = tiled. The tiled kernel (finetune path with full-length
     * gradients, nn/wav2vec2.py:437-444; 48 kHz pretraining) needs a caller-owned workspace of
     * a2v_attn_bwd_workspace_bytes() bytes, 16-byte aligned, and the call sequence
     *   a2v_attn_bwd_prepare (delta = rowsum(dO * O), clears the fp32 dQ accumulator)
     *   a2v_attn_bwd         (dK, dV written; dQ accumulated per key tile; qk_bound optional as in the forward)
     *   a2v_attn_bwd_finish  (dQ scaled and rounded into dqkv)
     * on one stream with the same descriptor. */
    int bwd_algo;
    void* workspace;
    int64_t workspace_bytes;
    /* backward, optional, bf16 resident kernel only (L <= 160): fp32 (3 * H * 64) vector that receives += the column sums
     * of dqkv, i.e. the gradient of the qkv Linear's bias (nn/modalities/modules.py:371), accumulated per CTA in shared
     * memory instead of by a separate pass over dqkv. Other kernels reject it (use a2v_colsum). */
    float* dqkv_colsum;
} a2v_attn_desc;

int a2v_attn_qk_bound(const void* qkv_bf16, float* bound, int batch, int L, int H, a2v_stream_t stream);
int a2v_attn_fwd(const a2v_attn_desc* d, a2v_stream_t stream);
size_t a2v_attn_bwd_workspace_bytes(int batch, int L, int H);
int a2v_attn_bwd_prepare(const a2v_attn_desc* d, a2v_stream_t stream);
int a2v_attn_bwd(const a2v_attn_desc* d, a2v_stream_t stream);
int a2v_attn_bwd_finish(const a2v_attn_desc* d, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * SincNet front end (nn/sinc.py:107-223,286-337). Channels-last output (B, N, 128) with the
 * 127 band-pass channels in columns 0..126 and a zero pad column. filters / dfilters are
 * (128, K) fp32 (row 127 zero). n_ and window_ are the (K/2,) fp32 buffers SincConv builds
 * at construction (sinc.py:264-276). Only stride 1 / dilation 1 / reflect "same" padding.
 * ------------------------------------------------------------------------------------ */
int a2v_sinc_filters_fwd(const float* low_hz, const float* band_hz, const float* n_, const float* window_, int C,
                         int K, float min_low_hz, float min_band_hz, float sample_rate, float* filters,
                         a2v_stream_t stream);
int a2v_sinc_filters_bwd(const float* low_hz, const float* band_hz, const float* n_, const float* window_, int C,
                         int K, float min_low_hz, float min_band_hz, float sample_rate, const float* dfilters,
                         float* dlow_hz, float* dband_hz, a2v_stream_t stream);
int a2v_sinc_conv_fwd(int out_dtype, const float* x, const float* filters, void* y, int B, int N, int K,
                      a2v_stream_t stream);
int a2v_sinc_conv_wgrad(int dy_dtype, const float* x, const void* dy, float* dfilters, int B, int N, int K,
                        a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Channels-last im2col / col2im for the strided feature-extractor convs (nn/utils.py:1085-1090).
 * ------------------------------------------------------------------------------------ */
int a2v_im2col(int dtype, const void* x, void* col, int B, int Tin, int Tout, int C, int k, int stride, int pad,
               a2v_stream_t stream);
int a2v_col2im(int dtype, const void* dcol, void* dx, int B, int Tin, int Tout, int C, int k, int stride, int pad,
               a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * BC-learning mixup (nn/data2vec2.py:453-498,536-598). gain_db is (B, W), W = (N-n_fft)/hop+1.
 * aweight = 10^(A-weighting dB / 10) for the n_fft/2+1 rfft bins (the reference builds it in
 * float64, data2vec2.py:461-479; pass it rounded to fp32). p_out (B,) optionally receives p.
 * ------------------------------------------------------------------------------------ */
int a2v_mixup_gain(const float* x, const float* hann, const float* aweight, int B, int N, int n_fft, int hop,
                   float min_db, float* gain_db, a2v_stream_t stream);
int a2v_mixup_apply(const float* x, const int32_t* perm, const float* gain_db, int B, int N, int W, float r,
                    float* out, float* p_out, a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Finetune head and criterion (BASELINE configs[3]).
 *   a2v_layer_mean_head_fwd: x = mean of the top-k FFN outputs, logits = proj(x) (nn/wav2vec2.py:446-464).
 *     layers = DEVICE table of K row-major (rows, D) pointers; W (C, D) / bias (C) fp32; xmean (rows, D, same dtype,
 *     optional) keeps the averaged input for the backward; logits (rows, C) fp32.
 *   a2v_head_bwd: g (rows, D, optional) = dlogits W / K (what each averaged layer output receives),
 *     dW += dlogits^T xmean, db += colsum(dlogits)  (fp32 accumulators).
 *   a2v_focal_loss_fwd/_bwd: sigmoid focal loss on logits (nn/utils.py:971-1010; alpha < 0 disables the class
 *     weighting) with the target mixup of nn/wav2vec2.py:424-431 folded in (perm = partner clip per clip, r = mixing
 *     weight; perm NULL = plain targets). loss_sum (double, accumulated), optional unreduced loss / mixed targets,
 *     counters[5] = {tp, fp, tn, fn, n_correct} of nn/criterions.py:198-229 + nn/utils.py:925-969 (multi-label
 *     branch, sigmoid >= threshold against the int64-truncated target), accumulated. grad_out: device scalar or NULL.
 *   a2v_channel_mask: x[b, t, c] = 0 where chmask[b, c] (nn/modalities/base.py:470-484), in place.
 * ------------------------------------------------------------------------------------ */
int a2v_layer_mean_head_fwd(int dtype, const void* const* layers, int K, int64_t rows, int D, int C, const float* W,
                            const float* bias, void* xmean, float* logits, a2v_stream_t stream);
int a2v_head_bwd(int dtype, const float* dlogits, const void* xmean, const float* W, int K, int64_t rows, int D, int C,
                 void* g, float* dW, float* db, a2v_stream_t stream);
int a2v_focal_loss_fwd(const float* logits, const float* targets, const int32_t* perm, int64_t rows, int C,
                       int rows_per_clip, float r, float alpha, float gamma, float threshold, double* loss_sum,
                       float* loss_out, float* mixed_targets, uint64_t* counters, a2v_stream_t stream);
int a2v_focal_loss_bwd(const float* logits, const float* targets, const int32_t* perm, int64_t rows, int C,
                       int rows_per_clip, float r, float alpha, float gamma, const float* grad_out, float* dlogits,
                       a2v_stream_t stream);
int a2v_channel_mask(int dtype, void* x, const uint8_t* chmask, int64_t rows, int rows_per_clip, int D,
                     a2v_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Input pipeline on the device: per-clip layer norm of a collated batch (fairseq
 * RawAudioDataset.postprocess with task.normalize, called at nn/audio_tasks.py:332) and frame-level multi-hot targets
 * from label intervals (nn/audio_tasks.py:336-381; intervals of clip b are [offsets[b], offsets[b+1]) in the flat
 * start / end / cat / foc arrays, sample positions round(linspace(0, wav_len, T, endpoint=False)); focal_class = index of
 * the "focal" class or -1).
 * ------------------------------------------------------------------------------------ */
int a2v_clip_layer_norm(const float* x, float* y, int B, int N, float eps, a2v_stream_t stream);
int a2v_frame_labels(const int32_t* offsets, const int32_t* start, const int32_t* end, const int32_t* cat,
                     const int32_t* foc, int B, int T, int C, int wav_len, int focal_class, float* out,
                     a2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* A2V_CAPI_H */
