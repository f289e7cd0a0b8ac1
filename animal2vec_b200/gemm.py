"""Host-side launchers for the tcgen05 GEMM (``a2v_gemm`` in include/a2v_capi.h).

Everything that is a matrix product on the pretraining path goes through here:
linear layers, (grouped) stride-1 convolutions expressed as a tap loop over shifted
TMA tiles, and the weight-gradient reductions. Operands are bf16 tensors; outputs are
bf16 or fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L


def _operand(t: torch.Tensor, dim0: int, dim1: int, dim2: int, stride1: int, stride2: int) -> L.Operand:
    if t.dtype != torch.bfloat16:
        raise L.A2VError(f"GEMM operands must be bfloat16, got {t.dtype}")
    return L.Operand(t.data_ptr(), dim0, dim1, dim2, stride1, stride2)


def _pick_block_n(n: int) -> int:
    if n <= 64:
        return 64
    if n <= 128:
        return 128
    return 256 if n % 256 == 0 or n > 512 else 128


def _flops(d: L.GemmDesc) -> float:
    if d.mode == 0:
        return 2.0 * d.M * d.N * d.k_per_tap * d.taps * d.batch * d.groups
    return 2.0 * d.M * d.N * d.red_rows * d.batch * d.taps * d.groups


def _launch(d: L.GemmDesc, anchor: torch.Tensor) -> None:
    L.require_device(anchor)
    L.launch_count += 1
    tl = L.gemm_timeline
    if L.op_timeline is not None:
        tag = (f"a2v_gemm[mode={d.mode},M={d.M},N={d.N},K={d.k_per_tap if d.mode == 0 else d.red_rows},taps={d.taps},"
               f"G={d.groups},B={d.batch},epi={int(bool(d.bias))}{int(d.act)}{int(bool(d.preact))}{int(bool(d.dgelu_u))}"
               f"{int(bool(d.residual))},tapM={d.a_tap_cols}]")
        L.timed_call(tag, lambda: L.check(L.load().a2v_gemm(C.byref(d), L.stream_ptr()), "a2v_gemm"))
        return
    if tl is None:
        L.check(L.load().a2v_gemm(C.byref(d), L.stream_ptr()), "a2v_gemm")
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.load().a2v_gemm(C.byref(d), L.stream_ptr()), "a2v_gemm")
    e1.record()
    kind = "linear" if (d.taps == 1 and d.groups == 1 and d.batch == 1) else "tap"
    tl.append((e0, e1, _flops(d), kind))


def _rows2d(t: torch.Tensor) -> tuple[int, int, int]:
    """(rows, cols, row_stride) of a tensor viewed as a row-major matrix over its last dim."""
    cols = t.shape[-1]
    rows = t.numel() // cols
    if t.dim() >= 2:
        t2 = t.reshape(rows, cols) if t.is_contiguous() else t
        if t2.dim() != 2:
            raise L.A2VError("non-contiguous operand with more than 2 dims")
        if t2.stride(1) != 1:
            raise L.A2VError("operand inner dimension must be contiguous")
        return rows, cols, t2.stride(0)
    return rows, cols, cols


def gemm_nt(
    a: torch.Tensor,
    w: torch.Tensor,
    *,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    bias: Optional[torch.Tensor] = None,
    act: int = 0,
    preact: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    dgelu_u: Optional[torch.Tensor] = None,
    colsum: Optional[torch.Tensor] = None,
    alpha: float = 1.0,
    accumulate: bool = False,
    block_n: Optional[int] = None,
) -> torch.Tensor:
    """out[m, n] = alpha * sum_k a[m, k] * w[n, k] (+bias[n]) -> GELU -> *GELU'(u) -> +residual.

    ``a``: (..., K) bf16, ``w``: (N, K) bf16 (nn.Linear weight layout). ``colsum`` (fp32 (N), with ``dgelu_u`` on the
    shapes :func:`pair_shape_ok` accepts): += column sums of the result.
    """
    m, k, lda = _rows2d(a)
    n, kw, ldb = _rows2d(w)
    if kw != k:
        raise L.A2VError(f"gemm_nt: K mismatch {k} vs {kw}")
    if out is None:
        out = torch.empty(*a.shape[:-1], n, device=a.device, dtype=out_dtype or torch.bfloat16)
    _, nc, ldc = _rows2d(out)
    assert nc == n
    for aux in (preact, residual, dgelu_u):
        if aux is not None:
            assert aux.dtype == out.dtype and aux.shape == out.shape and aux.is_contiguous() and out.is_contiguous()
    d = L.GemmDesc()
    d.mode = 0
    d.block_n = block_n or _pick_block_n(n)
    d.a = _operand(a, k, m, 1, lda, 0)
    d.b = _operand(w, k, n, 1, ldb, 0)
    d.M, d.N, d.k_per_tap, d.taps, d.batch, d.groups = m, n, k, 1, 1, 1
    d.k_splits = 1
    d.c = out.data_ptr()
    d.c_dtype = L.dtype_code(out)
    d.out_accumulate = 1 if accumulate else 0
    d.ldc = ldc
    d.alpha = alpha
    d.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
    d.act = act
    d.preact = preact.data_ptr() if preact is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.dgelu_u = dgelu_u.data_ptr() if dgelu_u is not None else None
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == n and colsum.is_contiguous() and dgelu_u is not None
        d.colsum = colsum.data_ptr()
    _launch(d, a)
    return out


def pair_shape_ok(m: int, n: int, k: int, dtype: torch.dtype = torch.bfloat16) -> bool:
    """Plain Linear shapes the CTA-pair kernels (gemm2_sm100.cu) take; the fused column-sum epilogue exists only there."""
    import os
    return (dtype == torch.bfloat16 and n % 256 == 0 and k % 64 == 0 and m >= 512 and
            os.environ.get("A2V_GEMM_2CTA", "1")[:1] != "0")


def conv_nt(
    x: torch.Tensor,
    w: torch.Tensor,
    *,
    taps: int,
    pad: int,
    groups: int = 1,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    bias: Optional[torch.Tensor] = None,
    accumulate: bool = False,
    block_n: Optional[int] = None,
) -> torch.Tensor:
    """Stride-1 (grouped) conv1d over channels-last activations without im2col.

    ``x``: (B, T, G*Cg) bf16; ``w``: (G*Ng, taps*Cg) bf16 with the tap index major and the
    in-group channel minor; returns (B, T, G*Ng):
        y[b, t, g*Ng + n] = sum_{j, c} x[b, t + j - pad, g*Cg + c] * w[g*Ng + n, j*Cg + c].
    Rows outside [0, T) read as zeros (TMA bounds check) = zero padding.
    """
    bsz, t, cin = x.shape
    assert x.is_contiguous() and w.is_contiguous()
    cg = cin // groups
    nout = w.shape[0]
    ng = nout // groups
    assert w.shape[1] == taps * cg, (w.shape, taps, cg)
    assert cg % 64 == 0 or taps == 1
    if out is None:
        out = torch.empty(bsz, t, nout, device=x.device, dtype=out_dtype or torch.bfloat16)
    assert out.is_contiguous()
    d = L.GemmDesc()
    d.mode = 0
    d.block_n = block_n or _pick_block_n(ng)
    d.a = _operand(x, cin, t, bsz, cin, t * cin)
    d.b = _operand(w, taps * cg, nout, 1, taps * cg, 0)
    d.M, d.N, d.k_per_tap, d.taps, d.batch, d.groups = t, ng, cg, taps, bsz, groups
    d.a_group_stride, d.a_row_off, d.a_tap_rows = cg, -pad, 1
    d.b_group_stride = ng
    d.k_splits = 1
    d.c = out.data_ptr()
    d.c_dtype = L.dtype_code(out)
    d.out_accumulate = 1 if accumulate else 0
    d.ldc = nout
    d.c_batch_stride = t
    d.c_group_stride = ng
    d.alpha = 1.0
    d.bias = bias.data_ptr() if bias is not None else None
    _launch(d, x)
    return out


def strided_conv_ok(t_in: int, c: int, k: int, stride: int, pad: int) -> bool:
    """The im2col-free strided path: T divisible by the stride, output length T / stride, 64-multiple channels."""
    return stride > 1 and t_in % stride == 0 and (t_in + 2 * pad - k) // stride + 1 == t_in // stride and c % 64 == 0


def strided_conv_nt(x: torch.Tensor, w: torch.Tensor, *, taps: int, stride: int, pad: int,
                    out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Strided Conv1d (nn/utils.py:1085-1092) over channels-last activations WITHOUT im2col.

    ``x``: (B, T, C) bf16, ``w``: (N, taps*C) bf16 tap-major (params.pack_conv_fwd); returns (B, T/stride, N):
        y[b, t, n] = sum_{j, c} x[b, stride*t + j - pad, c] * w[n, j*C + c]        (zero padding).
    x is addressed as (B, T/stride, stride*C): input row stride*t + q = (row t + floor(q/stride), column block q mod stride).
    """
    bsz, t, c = x.shape
    assert x.is_contiguous() and w.is_contiguous() and strided_conv_ok(t, c, taps, stride, pad)
    n = w.shape[0]
    assert w.shape[1] == taps * c
    tout = t // stride
    out = torch.empty(bsz, tout, n, device=x.device, dtype=out_dtype or torch.bfloat16)
    d = L.GemmDesc()
    d.mode = 0
    d.block_n = _pick_block_n(n)
    d.a = _operand(x, stride * c, tout, bsz, stride * c, t * c)
    d.b = _operand(w, taps * c, n, 1, taps * c, 0)
    d.M, d.N, d.k_per_tap, d.taps, d.batch, d.groups = tout, n, c, taps, bsz, 1
    d.a_group_stride, d.a_row_off, d.a_tap_rows = 0, -pad, 0
    d.a_tap_cols, d.a_tap_wrap = c, stride
    d.k_splits = 1
    d.c = out.data_ptr()
    d.c_dtype = L.dtype_code(out)
    d.ldc = n
    d.c_batch_stride = tout
    d.alpha = 1.0
    _launch(d, x)
    return out


def strided_conv_dgrad(dy: torch.Tensor, wd: torch.Tensor, *, c: int, taps_per_block: int, stride: int, pad: int,
                       out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Data gradient of :func:`strided_conv_nt` without col2im. ``dy``: (B, T/stride, N); ``wd``: (stride*C,
    taps_per_block*N) from params.pack_conv_dgrad_strided (block r = input rows congruent r mod stride, its taps
    j = (r + pad) mod stride + stride*u, zero weights where j >= k); returns dx (B, T, C):
        dx[b, stride*m + r, c] = sum_u sum_n dy[b, m + (r + pad)//stride - u, n] * w[n, c, (r + pad) % stride + stride*u].
    """
    bsz, tout, n = dy.shape
    assert dy.is_contiguous() and wd.is_contiguous() and wd.shape == (stride * c, taps_per_block * n) and n % 64 == 0
    dx = torch.empty(bsz, tout * stride, c, device=dy.device, dtype=out_dtype or torch.bfloat16)
    d = L.GemmDesc()
    d.mode = 0
    d.block_n = _pick_block_n(c)
    d.a = _operand(dy, n, tout, bsz, n, tout * n)
    d.b = _operand(wd, taps_per_block * n, stride * c, 1, taps_per_block * n, 0)
    d.M, d.N, d.k_per_tap, d.taps, d.batch, d.groups = tout, c, n, taps_per_block, bsz, stride
    d.a_group_stride, d.a_row_off, d.a_tap_rows = 0, 0, -1
    d.a_grow_add, d.a_grow_div = pad, stride
    d.b_group_stride = c
    d.k_splits = 1
    d.c = dx.data_ptr()
    d.c_dtype = L.dtype_code(dx)
    d.ldc = stride * c
    d.c_batch_stride = tout
    d.c_group_stride = c
    d.alpha = 1.0
    _launch(d, dy)
    return dx


def strided_conv_wgrad_tn(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, *, taps: int, stride: int, pad: int,
                          k_splits: Optional[int] = None) -> torch.Tensor:
    """Weight gradient of :func:`strided_conv_nt`, atomically accumulated into fp32 ``out`` (taps*C, N) (the
    transposed tap-major layout of conv_wgrad_tn): out[j*C + c, n] += sum_{b, t} x[b, stride*t + j - pad, c] * dy[b, t, n]."""
    bsz, t, c = x.shape
    tout, n = dy.shape[1], dy.shape[2]
    assert dy.is_contiguous() and x.is_contiguous() and tout == t // stride and strided_conv_ok(t, c, taps, stride, pad)
    assert out.dtype == torch.float32 and out.shape == (taps * c, n) and out.is_contiguous()
    d = L.GemmDesc()
    d.mode = 1
    d.block_n = 64 if n <= 64 else (128 if n <= 128 else 256)
    d.a = _operand(x, stride * c, tout, bsz, stride * c, t * c)
    d.b = _operand(dy, n, tout, bsz, n, tout * n)
    d.M, d.N, d.taps, d.batch, d.groups = taps * c, n, 1, bsz, 1
    d.a_group_stride, d.a_row_off, d.a_tap_rows, d.a_tap_cols, d.a_tap_wrap = 0, -pad, 0, c, stride
    d.b_group_stride, d.b_row_off, d.b_tap_rows = 0, 0, 0
    d.red_rows = tout
    tiles = -(-(taps * c) // 128) * -(-n // d.block_n)
    kblocks = bsz * -(-tout // 64)
    ks = k_splits or _pick_splits(tiles, kblocks)
    per = -(-kblocks // ks)
    d.k_splits = -(-kblocks // per)
    d.c = out.data_ptr()
    d.c_dtype = L.F32
    d.out_atomic = 1
    d.ldc = n
    d.c_group_stride = taps * c
    d.c_tap_stride = 0
    d.alpha = 1.0
    _launch(d, x)
    return out


def gathered_conv_nt(xg: torch.Tensor, w: torch.Tensor, *, taps: int, groups: int, bias: Optional[torch.Tensor] = None,
                     out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Grouped conv evaluated only at pre-gathered positions. ``xg``: (M, taps, C) bf16 (row m, tap j = input frame
    pos_m + j - pad, zeros outside the sequence), ``w``: (G*Ng, taps*Cg) tap-major (params.pack_conv_fwd):
        y[m, g*Ng + n] = bias + sum_{j, c} xg[m, j, g*Cg + c] * w[g*Ng + n, j*Cg + c]."""
    m, taps_, c = xg.shape
    assert taps_ == taps and xg.is_contiguous() and w.is_contiguous()
    cg = c // groups
    nout = w.shape[0]
    ng = nout // groups
    assert w.shape[1] == taps * cg and cg % 64 == 0
    out = torch.empty(m, nout, device=xg.device, dtype=out_dtype or torch.bfloat16)
    d = L.GemmDesc()
    d.mode = 0
    d.block_n = _pick_block_n(ng)
    d.a = _operand(xg, taps * c, m, 1, taps * c, 0)
    d.b = _operand(w, taps * cg, nout, 1, taps * cg, 0)
    d.M, d.N, d.k_per_tap, d.taps, d.batch, d.groups = m, ng, cg, taps, 1, groups
    d.a_group_stride, d.a_row_off, d.a_tap_rows = cg, 0, 0
    d.a_tap_cols, d.a_tap_wrap, d.a_tap_col_stride = cg, taps, c
    d.b_group_stride = ng
    d.k_splits = 1
    d.c = out.data_ptr()
    d.c_dtype = L.dtype_code(out)
    d.ldc = nout
    d.c_group_stride = ng
    d.alpha = 1.0
    d.bias = bias.data_ptr() if bias is not None else None
    _launch(d, xg)
    return out


def gathered_conv_wgrad_tn(dy: torch.Tensor, xg: torch.Tensor, out: torch.Tensor, *, taps: int, groups: int,
                           k_splits: Optional[int] = None) -> torch.Tensor:
    """Weight gradient of :func:`gathered_conv_nt` in the transposed tap-major layout of conv_wgrad_tn:
    out[(g*taps + j)*Cg + c, n] += sum_m xg[m, j, g*Cg + c] * dy[m, g*Ng + n]   (fp32, atomically accumulated)."""
    m, taps_, c = xg.shape
    nout = dy.shape[-1]
    cg, ng = c // groups, nout // groups
    assert taps_ == taps and dy.shape[0] == m and dy.is_contiguous() and xg.is_contiguous() and cg % 64 == 0
    assert out.dtype == torch.float32 and out.shape == (groups * taps * cg, ng) and out.is_contiguous()
    d = L.GemmDesc()
    d.mode = 1
    d.block_n = 64 if ng <= 64 else (128 if ng <= 128 else 256)
    d.a = _operand(xg, taps * c, m, 1, taps * c, 0)
    d.b = _operand(dy, nout, m, 1, nout, 0)
    d.M, d.N, d.taps, d.batch, d.groups = taps * cg, ng, 1, 1, groups
    d.a_group_stride, d.a_row_off, d.a_tap_rows = cg, 0, 0
    d.a_tap_cols, d.a_tap_wrap, d.a_tap_col_stride = cg, taps, c
    d.b_group_stride, d.b_row_off, d.b_tap_rows = ng, 0, 0
    d.red_rows = m
    tiles = -(-(taps * cg) // 128) * -(-ng // d.block_n) * groups
    kblocks = -(-m // 64)
    ks = k_splits or (_pick_splits(tiles, kblocks) if tiles < 148 else _pick_splits_waves(tiles, kblocks))
    per = -(-kblocks // ks)
    d.k_splits = -(-kblocks // per)
    d.c = out.data_ptr()
    d.c_dtype = L.F32
    d.out_atomic = 1
    d.ldc = ng
    d.c_group_stride = taps * cg
    d.c_tap_stride = 0
    d.alpha = 1.0
    _launch(d, xg)
    return out


def conv_slab_ok(x: torch.Tensor, w: torch.Tensor, taps: int, groups: int) -> bool:
    """The slab kernel covers bf16 tap convs with 64-channel groups (or compact 48-of-64 groups, see :func:`conv_slab`)
    and <= 64 outputs per group."""
    if x.dtype != torch.bfloat16 or x.dim() != 3 or taps > 32:
        return False
    cin, nout = x.shape[-1], w.shape[0]
    return cin in (groups * 64, groups * 48) and nout // groups <= 64 and w.shape[1] == taps * 64


def conv_slab(x: torch.Tensor, w: torch.Tensor, *, taps: int, pad: int, groups: int,
              out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
              bias: Optional[torch.Tensor] = None, x_real_cols: int = 0, ng_out: Optional[int] = None) -> torch.Tensor:
    """Same operator as :func:`conv_nt` (64-channel groups), every activation row read once per tile. ``x_real_cols``:
    channels per 64-wide input group that can be non-zero (group-padded layouts); the K steps over the rest are skipped.
    Compact group-padded layouts: ``x`` may store its groups ``x_real_cols`` = 48 channels apart (x.shape[-1] =
    groups * 48; the weights keep 64 K columns per tap), and ``ng_out`` < w.shape[0] // groups writes only the first
    ``ng_out`` outputs of every group, ``ng_out`` columns apart (the weight rows beyond are padding)."""
    bsz, t, cin = x.shape
    assert x.is_contiguous() and w.is_contiguous() and w.dtype == torch.bfloat16
    wrows = w.shape[0] // groups
    ng = ng_out or wrows
    nout = groups * ng
    xg = cin // groups
    assert xg == 64 or (xg == x_real_cols and xg % 16 == 0), (xg, x_real_cols)
    if out is None:
        out = torch.empty(bsz, t, nout, device=x.device, dtype=out_dtype or torch.bfloat16)
    assert out.is_contiguous() and out.shape == (bsz, t, nout)
    d = L.ConvDesc()
    d.x, d.w, d.y = x.data_ptr(), w.data_ptr(), out.data_ptr()
    d.batch, d.T, d.groups, d.taps, d.pad = bsz, t, groups, taps, pad
    d.ng, d.x_group_cols, d.w_group_rows, d.y_group_cols = ng, xg, wrows, ng
    d.ldx, d.ldw, d.ldy = cin, w.shape[1], nout
    d.y_dtype = L.dtype_code(out)
    d.bias = bias.data_ptr() if bias is not None else None
    d.x_real_cols = x_real_cols
    L.require_device(x)
    L.launch_count += 1
    tl = L.gemm_timeline
    if L.op_timeline is not None:
        L.timed_call(f"a2v_conv_slab_fwd[B={bsz},T={t},G={groups},taps={taps},ng={ng}]",
                     lambda: L.check(L.load().a2v_conv_slab_fwd(C.byref(d), L.stream_ptr()), "a2v_conv_slab_fwd"))
        return out
    if tl is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(L.load().a2v_conv_slab_fwd(C.byref(d), L.stream_ptr()), "a2v_conv_slab_fwd")
    if tl is not None:
        e1.record()
        tl.append((e0, e1, 2.0 * bsz * t * nout * (x_real_cols or 64) * taps, "slab"))
    return out


def conv_slab_wgrad(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, *, taps: int, pad: int, groups: int) -> torch.Tensor:
    """Same contract as :func:`conv_wgrad_tn` (transposed (G*taps*64, Ng) fp32 output) for 64-channel groups:
    one x slab per 64-row k-block serves all taps of a 16-tap block. Compact group-padded layouts: ``x`` with its groups
    48 channels apart (rows c >= 48 of every 64-row tap block of ``out`` are scratch) and / or ``dy`` with fewer columns
    per group than ``out`` has (the columns beyond are left untouched)."""
    bsz, t, cin = x.shape
    nout = dy.shape[-1]
    ng = nout // groups
    xg = cin // groups
    assert xg in (64, 48) and ng <= 64 and dy.shape[:2] == x.shape[:2] and dy.is_contiguous() and x.is_contiguous()
    assert out.dtype == torch.float32 and out.shape[0] == groups * taps * 64 and out.shape[1] >= ng and out.is_contiguous()
    d = L.ConvDesc()
    d.x, d.w, d.y = x.data_ptr(), dy.data_ptr(), None
    d.batch, d.T, d.groups, d.taps, d.pad = bsz, t, groups, taps, pad
    d.ng, d.x_group_cols, d.w_group_rows, d.y_group_cols = ng, xg, ng, ng
    d.x_real_cols = xg if xg < 64 else 0
    d.ldx, d.ldw, d.ldy = cin, nout, out.shape[1]
    d.y_dtype = L.F32
    L.require_device(x)
    L.launch_count += 1
    fn = lambda: L.check(L.load().a2v_conv_slab_wgrad(C.byref(d), C.c_void_p(out.data_ptr()), C.c_int64(out.shape[1]),
                                                      L.stream_ptr()), "a2v_conv_slab_wgrad")
    tl = L.gemm_timeline
    if L.op_timeline is not None:
        L.timed_call(f"a2v_conv_slab_wgrad[B={bsz},T={t},G={groups},taps={taps},ng={ng}]", fn)
    elif tl is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        tl.append((e0, e1, 2.0 * bsz * t * nout * xg * taps, "slab"))
    else:
        fn()
    return out


def _pick_splits_waves(tiles: int, kblocks: int, max_splits: int = 8) -> int:
    """Split count for tile counts slightly above the SM count: the split that wastes the least of its last wave
    (160 tiles on 148 SMs run as two waves with the second 8 % full; 7 splits fill 7.6 of 8 waves)."""
    sms = 148
    best, best_eff = 1, 0.0
    for s_ in range(1, max_splits + 1):
        if kblocks // s_ < 8:
            break
        work = tiles * s_
        eff = work / (-(-work // sms) * sms)
        if eff > best_eff + 0.02:
            best, best_eff = s_, eff
    return best


def _pick_splits(tiles: int, kblocks: int) -> int:
    sms = 148
    if tiles >= sms:
        return 1
    want = max(1, (2 * sms) // max(tiles, 1))
    want = min(want, max(1, kblocks // 4))
    per = -(-kblocks // want)
    return -(-kblocks // per)


def gemm_tn(
    a: torch.Tensor,
    b: torch.Tensor,
    out: torch.Tensor,
    *,
    alpha: float = 1.0,
    k_splits: Optional[int] = None,
    block_n: Optional[int] = None,
) -> torch.Tensor:
    """out[m, n] += alpha * sum_r a[r, m] * b[r, n]  (fp32 ``out``, atomically accumulated).

    ``a``: (R, M) bf16, ``b``: (R, N) bf16 -- the weight-gradient product dY^T X.
    """
    r, m, lda = _rows2d(a)
    r2, n, ldb = _rows2d(b)
    assert r == r2 and out.dtype == torch.float32 and out.is_contiguous()
    assert out.shape[-2:] == (m, n) or out.numel() == m * n
    d = L.GemmDesc()
    d.mode = 1
    d.block_n = block_n or _pick_block_n(n)
    d.a = _operand(a, m, r, 1, lda, 0)
    d.b = _operand(b, n, r, 1, ldb, 0)
    d.M, d.N, d.taps, d.batch, d.groups = m, n, 1, 1, 1
    d.red_rows = r
    tiles = -(-m // 128) * -(-n // d.block_n)
    kblocks = -(-r // 64)
    ks = k_splits or _pick_splits(tiles, kblocks)
    per = -(-kblocks // ks)
    d.k_splits = -(-kblocks // per)
    d.c = out.data_ptr()
    d.c_dtype = L.F32
    d.out_atomic = 1
    d.ldc = n
    d.alpha = alpha
    _launch(d, a)
    return out


def conv_wgrad_tn(
    dy: torch.Tensor,
    x: torch.Tensor,
    out: torch.Tensor,
    *,
    taps: int,
    pad: int,
    groups: int = 1,
    k_splits: Optional[int] = None,
) -> torch.Tensor:
    """Weight gradient of :func:`conv_nt`, atomically accumulated into fp32 ``out`` of shape
    (G*taps*Cg, Ng) -- TRANSPOSED relative to the forward weight layout so that x (with the taps folded
    into M, two taps per 128-row MMA tile) is the M side and the tile rows are contiguous in ``out``:
        out[(g*taps + j)*Cg + c, n] += sum_{b, t} x[b, t + j - pad, g*Cg + c] * dy[b, t, g*Ng + n].
    """
    bsz, t, cin = x.shape
    nout = dy.shape[-1]
    cg, ng = cin // groups, nout // groups
    assert dy.shape[:2] == x.shape[:2] and dy.is_contiguous() and x.is_contiguous()
    assert cg % 64 == 0, "conv_wgrad_tn: channels per group must be a multiple of 64"
    assert out.dtype == torch.float32 and out.shape == (groups * taps * cg, ng) and out.is_contiguous()
    d = L.GemmDesc()
    d.mode = 1
    d.block_n = 64 if ng <= 64 else (128 if ng <= 128 else 256)
    d.a = _operand(x, cin, t, bsz, cin, t * cin)
    d.b = _operand(dy, nout, t, bsz, nout, t * nout)
    d.M, d.N, d.taps, d.batch, d.groups = taps * cg, ng, 1, bsz, groups
    d.a_group_stride, d.a_row_off, d.a_tap_rows, d.a_tap_cols = cg, -pad, 1, cg
    d.b_group_stride, d.b_row_off, d.b_tap_rows = ng, 0, 0
    d.red_rows = t
    tiles = -(-(taps * cg) // 128) * -(-ng // d.block_n) * groups
    kblocks = bsz * -(-t // 64)
    ks = k_splits or _pick_splits(tiles, kblocks)
    per = -(-kblocks // ks)
    d.k_splits = -(-kblocks // per)
    d.c = out.data_ptr()
    d.c_dtype = L.F32
    d.out_atomic = 1
    d.ldc = ng
    d.c_group_stride = taps * cg
    d.c_tap_stride = 0
    d.alpha = 1.0
    _launch(d, x)
    return out
