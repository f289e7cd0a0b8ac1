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
} a2v_gemm_desc;

int a2v_gemm(const a2v_gemm_desc* d, a2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* A2V_CAPI_H */
