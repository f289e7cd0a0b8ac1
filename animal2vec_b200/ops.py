"""Thin, typed Python launchers over the C-ABI (one function per entry point of
include/a2v_capi.h, GEMMs are in :mod:`animal2vec_b200.gemm`). No arithmetic happens here:
these functions only allocate outputs with torch and pass raw pointers + the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, List

import torch

from . import lib as L


class RowLnDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("rows", C.c_int64),
        ("channels", C.c_int),
        ("group_width", C.c_int),
        ("group_real", C.c_int),
        ("eps", C.c_float),
        ("act", C.c_int),
        ("a", C.c_void_p),
        ("b", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("act_alpha", C.c_void_p),
        ("act_beta", C.c_void_p),
        ("post", C.c_void_p),
        ("y", C.c_void_p),
        ("mean", C.c_void_p),
        ("rstd", C.c_void_p),
        ("drop_b", C.c_float),
        ("seed_b", C.c_uint64),
        ("drop_out", C.c_float),
        ("seed_out", C.c_uint64),
        ("dy", C.c_void_p),
        ("da", C.c_void_p),
        ("db", C.c_void_p),
        ("dgamma", C.c_void_p),
        ("dbeta", C.c_void_p),
        ("dact_alpha", C.c_void_p),
        ("dact_beta", C.c_void_p),
        ("dbias_b", C.c_void_p),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("batch", C.c_int),
        ("L", C.c_int),
        ("H", C.c_int),
        ("head_dim", C.c_int),
        ("qkv", C.c_void_p),
        ("out", C.c_void_p),
        ("lse", C.c_void_p),
        ("pos", C.c_void_p),
        ("slopes", C.c_void_p),
        ("alibi_scale", C.c_void_p),
        ("alibi_scale_stride", C.c_int),
        ("sm_scale", C.c_float),
        ("drop_p", C.c_float),
        ("seed", C.c_uint64),
        ("dout", C.c_void_p),
        ("dqkv", C.c_void_p),
        ("dalibi_scale", C.c_void_p),
        ("qk_bound", C.c_void_p),
        ("bwd_algo", C.c_int),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("dqkv_colsum", C.c_void_p),
    ]


def _p(t: Optional[torch.Tensor]):
    """Raw device pointer as a ctypes void* (None -> NULL). Plain ints would be truncated to 32 bits."""
    return None if t is None else C.c_void_p(t.data_ptr())


# Host -> device staging without a stream synchronisation. A copy from PAGEABLE host memory makes the
# runtime synchronise the stream first, which stalls the launch queue once per call; small per-step host
# values (masks, permutations, pointer tables) therefore travel through a ring of pinned buffers.
_h2d_rings: dict = {}


def h2d_async(src: torch.Tensor, device, slots: int = 4) -> torch.Tensor:
    """Device copy of a CPU tensor via pinned staging (stream-ordered, never blocks on the GPU)."""
    src = src.contiguous()
    key = (src.dtype, src.numel(), str(device))
    ring = _h2d_rings.get(key)
    if ring is None:
        ring = _h2d_rings[key] = {"buf": [None] * slots, "ev": [None] * slots, "i": 0}
    i = ring["i"]
    ring["i"] = (i + 1) % slots
    if ring["buf"][i] is None:
        ring["buf"][i] = torch.empty(src.numel(), dtype=src.dtype).pin_memory()
    else:
        ring["ev"][i].synchronize()  # the copy that last used this slot (several steps ago) has completed
    buf = ring["buf"][i]
    buf.copy_(src.view(-1))
    out = buf.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    ring["ev"][i] = ev
    return out.view(src.shape)


def _call(name: str, anchor: torch.Tensor, *args) -> None:
    L.require_device(anchor)
    fn = getattr(L.load(), name)
    L.launch_count += 1
    if L.op_timeline is not None:
        tag = name
        if name.startswith("a2v_rowln"):
            d = args[0]._obj
            tag = f"{name}[C={d.channels},act={d.act},aff={int(bool(d.gamma))},b={int(bool(d.b))},rows={d.rows}]"
        elif name in ("a2v_attn_fwd", "a2v_attn_bwd"):
            d = args[0]._obj
            tag = f"{name}[L={d.L},batch={d.batch}]"
        L.timed_call(tag, lambda: L.check(fn(*args, L.stream_ptr()), name))
        return
    L.check(fn(*args, L.stream_ptr()), name)


# ----------------------------------------------------------------------------- row LN
class RowLnCfg:
    """Static description of one fused LN site (see a2v_rowln_desc)."""

    __slots__ = ("channels", "group_width", "group_real", "eps", "act", "drop_b", "drop_out")

    def __init__(self, channels, eps, act=0, group_width=None, group_real=None, drop_b=0.0, drop_out=0.0):
        self.channels = channels
        self.group_width = group_width or channels
        self.group_real = group_real or self.group_width
        self.eps = eps
        self.act = act
        self.drop_b = drop_b
        self.drop_out = drop_out


def _rowln_desc(cfg: RowLnCfg, a, b, gamma, beta, act_alpha, act_beta, post, y, mean, rstd, seed_b, seed_out,
                training: bool) -> RowLnDesc:
    d = RowLnDesc()
    d.dtype = L.dtype_code(a)
    d.rows = a.numel() // cfg.channels
    d.channels, d.group_width, d.group_real = cfg.channels, cfg.group_width, cfg.group_real
    d.eps, d.act = cfg.eps, cfg.act
    d.a, d.b, d.gamma, d.beta = _p(a), _p(b), _p(gamma), _p(beta)
    d.act_alpha, d.act_beta, d.post = _p(act_alpha), _p(act_beta), _p(post)
    d.y, d.mean, d.rstd = _p(y), _p(mean), _p(rstd)
    d.drop_b = cfg.drop_b if training else 0.0
    d.drop_out = cfg.drop_out if training else 0.0
    d.seed_b, d.seed_out = seed_b, seed_out
    return d


def rowln_fwd(cfg: RowLnCfg, a, b=None, gamma=None, beta=None, act_alpha=None, act_beta=None, post=None,
              seed_b=0, seed_out=0, training=True, save_stats=True):
    assert a.is_contiguous() and a.shape[-1] == cfg.channels
    for t in (b, post):
        assert t is None or (t.is_contiguous() and t.shape == a.shape and t.dtype == a.dtype)
    y = torch.empty_like(a)
    rows = a.numel() // cfg.channels
    mean = torch.empty(rows, device=a.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=a.device, dtype=torch.float32) if save_stats else None
    d = _rowln_desc(cfg, a, b, gamma, beta, act_alpha, act_beta, post, y, mean, rstd, seed_b, seed_out, training)
    _call("a2v_rowln_fwd", a, C.byref(d))
    return y, mean, rstd


def rowln_bwd(cfg: RowLnCfg, dy, a, b, gamma, beta, act_alpha, act_beta, mean, rstd, *, seed_b=0, seed_out=0,
              training=True, want_db=True, dgamma=None, dbeta=None, dact_alpha=None, dact_beta=None, dbias_b=None):
    """Returns (da, db). Parameter gradients are accumulated into the given fp32 buffers; ``dbias_b`` (optional)
    accumulates the column sums of db, i.e. the bias gradient of the Linear whose output was ``b``."""
    assert dy.is_contiguous() and dy.dtype == a.dtype
    da = torch.empty_like(a)
    db = torch.empty_like(a) if (b is not None and want_db) else None
    d = _rowln_desc(cfg, a, b, gamma, beta, act_alpha, act_beta, None, None, mean, rstd, seed_b, seed_out, training)
    d.dy, d.da, d.db = _p(dy), _p(da), _p(db)
    d.dgamma, d.dbeta, d.dact_alpha, d.dact_beta = _p(dgamma), _p(dbeta), _p(dact_alpha), _p(dact_beta)
    d.dbias_b = _p(dbias_b)
    _call("a2v_rowln_bwd", a, C.byref(d))
    return da, db


# ----------------------------------------------------------------------------- attention
def attn_fwd(qkv, batch, seq, heads, *, pos=None, slopes=None, alibi_scale=None, drop_p=0.0, seed=0,
             need_lse=True, skip_far_keys=False):
    """``skip_far_keys`` (bf16, contiguous positions, ALiBi slopes given): key tiles whose every probability is provably
    below 2^-50 of the row maximum are not visited (one extra launch computes max|q| max|k| per head)."""
    assert qkv.is_contiguous() and qkv.shape[-1] == 3 * heads * 64
    out = torch.empty(batch, seq, heads * 64, device=qkv.device, dtype=qkv.dtype)
    bound = None
    if skip_far_keys and pos is None and slopes is not None and qkv.dtype == torch.bfloat16 and seq > 256:
        bound = torch.zeros(batch * heads * 2, device=qkv.device, dtype=torch.float32)
        _call("a2v_attn_qk_bound", qkv, _p(qkv), _p(bound), batch, seq, heads)
    lse = torch.empty(batch, heads, seq, device=qkv.device, dtype=torch.float32) if need_lse else None
    d = AttnDesc()
    d.dtype = L.dtype_code(qkv)
    d.batch, d.L, d.H, d.head_dim = batch, seq, heads, 64
    d.qkv, d.out, d.lse, d.pos = _p(qkv), _p(out), _p(lse), _p(pos)
    d.slopes, d.alibi_scale = _p(slopes), _p(alibi_scale)
    d.alibi_scale_stride = 0 if (alibi_scale is None or alibi_scale.numel() == 1) else 1
    d.sm_scale = 64 ** -0.5
    d.drop_p, d.seed = drop_p, seed
    d.qk_bound = _p(bound)
    _call("a2v_attn_fwd", qkv, C.byref(d))
    return out, lse


BWD_RESIDENT_MAX = 160  # tokens per sequence the shared-memory-resident backward takes (attention.cu BWD_LMAX)


def attn_bwd(dout, qkv, out, lse, batch, seq, heads, *, pos=None, slopes=None, alibi_scale=None,
             dalibi_scale=None, drop_p=0.0, seed=0, algo=0, skip_far_keys=True, dqkv_colsum=None):
    """``algo``: 0 automatic (bf16: resident kernel up to 160 tokens, tiled kernel beyond), 1 resident, 2 tiled.
    The tiled kernel runs as prepare -> backward -> finish (three launches) over a workspace allocated here."""
    assert dout.is_contiguous() and dout.dtype == qkv.dtype
    dqkv = torch.zeros_like(qkv) if qkv.dtype == torch.float32 else torch.empty_like(qkv)
    tiled = qkv.dtype == torch.bfloat16 and (algo == 2 or (algo == 0 and seq > BWD_RESIDENT_MAX))
    d = AttnDesc()
    d.bwd_algo = algo
    if tiled:
        lib = L.load()
        lib.a2v_attn_bwd_workspace_bytes.restype = C.c_size_t
        nbytes = int(lib.a2v_attn_bwd_workspace_bytes(batch, seq, heads))
        ws = torch.empty((nbytes + 3) // 4, device=qkv.device, dtype=torch.float32)
        d.workspace, d.workspace_bytes = _p(ws), nbytes
        if skip_far_keys and pos is None and slopes is not None and seq > 256:
            bound = torch.zeros(batch * heads * 2, device=qkv.device, dtype=torch.float32)
            _call("a2v_attn_qk_bound", qkv, _p(qkv), _p(bound), batch, seq, heads)
            d.qk_bound = _p(bound)
    d.dtype = L.dtype_code(qkv)
    d.batch, d.L, d.H, d.head_dim = batch, seq, heads, 64
    d.qkv, d.out, d.lse, d.pos = _p(qkv), _p(out), _p(lse), _p(pos)
    d.slopes, d.alibi_scale = _p(slopes), _p(alibi_scale)
    d.alibi_scale_stride = 0 if (alibi_scale is None or alibi_scale.numel() == 1) else 1
    d.sm_scale = 64 ** -0.5
    d.drop_p, d.seed = drop_p, seed
    d.dout, d.dqkv, d.dalibi_scale = _p(dout), _p(dqkv), _p(dalibi_scale)
    fuse_colsum = (dqkv_colsum is not None and not tiled and qkv.dtype == torch.bfloat16 and seq <= BWD_RESIDENT_MAX
                   and heads <= 128)
    if fuse_colsum:  # qkv-bias gradient accumulated inside the resident kernel
        assert dqkv_colsum.dtype == torch.float32 and dqkv_colsum.numel() == 3 * heads * 64
        d.dqkv_colsum = _p(dqkv_colsum)
    if tiled:
        _call("a2v_attn_bwd_prepare", qkv, C.byref(d))
        _call("a2v_attn_bwd", qkv, C.byref(d))
        _call("a2v_attn_bwd_finish", qkv, C.byref(d))
    else:
        _call("a2v_attn_bwd", qkv, C.byref(d))
    if dqkv_colsum is not None and not fuse_colsum:
        colsum(dqkv.view(-1, dqkv.shape[-1]), dqkv_colsum)
    return dqkv


# ----------------------------------------------------------------------------- masking
class MaskIndex:
    """Device-side MaskInfo (nn/modalities/base.py:427-455) plus the flat row maps."""

    __slots__ = ("mask", "ids_keep", "ids_restore", "clone_src", "keep_src_x", "keep_src_clone", "restore_src",
                 "err", "rows", "T", "Tk", "clones")


def mask_index(mask_u8: torch.Tensor, tk: int, clones: int) -> MaskIndex:
    rows, t = mask_u8.shape
    assert mask_u8.dtype == torch.uint8 and mask_u8.is_contiguous()
    dev = mask_u8.device
    mi = MaskIndex()
    mi.mask, mi.rows, mi.T, mi.Tk, mi.clones = mask_u8, rows, t, tk, clones
    i32 = dict(device=dev, dtype=torch.int32)
    mi.ids_keep = torch.empty(rows, tk, **i32)
    mi.ids_restore = torch.empty(rows, t, **i32)
    mi.clone_src = torch.empty(rows * t, **i32)
    mi.keep_src_x = torch.empty(rows * tk, **i32)
    mi.keep_src_clone = torch.empty(rows * tk, **i32)
    mi.restore_src = torch.empty(rows * t, **i32)
    mi.err = torch.zeros(1, **i32)
    _call("a2v_mask_index", mask_u8, _p(mask_u8), rows, t, tk, clones, _p(mi.ids_keep), _p(mi.ids_restore),
          _p(mi.clone_src), _p(mi.keep_src_x), _p(mi.keep_src_clone), _p(mi.restore_src), _p(mi.err))
    return mi


def row_gather(src, idx, n_dst, *, add=None, fill_std=0.0, fill_seed=0, drop_p=0.0, drop_seed=0,
               drop_by_src=False, out_shape=None):
    d_ = src.shape[-1]
    assert src.is_contiguous() and idx.dtype == torch.int32 and idx.numel() == n_dst
    dst = torch.empty(out_shape or (n_dst, d_), device=src.device, dtype=src.dtype)
    assert dst.numel() == n_dst * d_
    if add is not None:
        assert add.is_contiguous() and add.numel() == dst.numel() and add.dtype == src.dtype
    _call("a2v_row_gather", src, L.dtype_code(src), _p(src), _p(idx), _p(add), _p(dst), C.c_int64(n_dst), d_,
          C.c_float(fill_std), C.c_uint64(fill_seed), C.c_float(drop_p), C.c_uint64(drop_seed), int(drop_by_src))
    return dst


def neigh_index(ids_keep: torch.Tensor, t: int, taps: int, pad: int) -> torch.Tensor:
    """(rows*Tk*taps,) int32 row map: flat row r*T + ids_keep[r, i] + j - pad of the (rows*T, C) activation, -1 outside."""
    rows, tk = ids_keep.shape
    assert ids_keep.dtype == torch.int32 and ids_keep.is_contiguous()
    out = torch.empty(rows * tk * taps, device=ids_keep.device, dtype=torch.int32)
    _call("a2v_neigh_index", ids_keep, _p(ids_keep), rows, tk, t, taps, pad, _p(out))
    return out


def clone_sum_bwd(d_masked, d_unmasked, restore_src, b, t, clones, d_):
    anchor = d_masked if d_masked is not None else d_unmasked
    dx = torch.empty(b, t, d_, device=anchor.device, dtype=anchor.dtype)
    _call("a2v_clone_sum_bwd", anchor, L.dtype_code(anchor), _p(d_masked), _p(d_unmasked), _p(restore_src), _p(dx),
          C.c_int64(b), t, clones, d_)
    return dx


# ----------------------------------------------------------------------------- targets / loss
def make_targets(layers: Sequence[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    k = len(layers)
    b, t, d_ = layers[0].shape
    for x in layers:
        assert x.is_contiguous() and x.shape == layers[0].shape and x.dtype == layers[0].dtype
    dev = layers[0].device
    ptrs = h2d_async(torch.tensor([x.data_ptr() for x in layers], dtype=torch.int64), dev)
    stats = torch.empty(k, b, d_, 2, device=dev, dtype=torch.float32)
    y = torch.empty(b, t, d_, device=dev, dtype=torch.float32)
    code = L.dtype_code(layers[0])
    _call("a2v_target_stats", layers[0], code, _p(ptrs), k, b, t, d_, C.c_float(eps), _p(stats))
    _call("a2v_target_apply", layers[0], code, _p(ptrs), k, b, t, d_, _p(stats), _p(y))
    return y


def d2v_loss_fwd(pred, y, mask_u8, clones, scale):
    r, t, d_ = pred.shape
    assert pred.is_contiguous() and y.is_contiguous() and y.dtype == torch.float32
    acc = torch.zeros(1 + 4 * d_, device=pred.device, dtype=torch.float64)
    _call("a2v_d2v_loss_fwd", pred, L.dtype_code(pred), _p(pred), _p(y), _p(mask_u8), C.c_int64(r), t, clones, d_,
          C.c_float(scale), C.c_void_p(acc.data_ptr()), C.c_void_p(acc.data_ptr() + 8))
    return acc[0:1], acc[1:].view(4, d_)


def d2v_loss_fused(pred, y, mask_u8, clones, scale, grad_coef: Optional[float]):
    """One pass: (loss_sum, colstats) and -- when ``grad_coef`` is given -- the gradient grad_coef * (pred - y) on masked
    rows, zeros elsewhere, written IN PLACE over ``pred`` (bf16, D % 256 == 0)."""
    r, t, d_ = pred.shape
    assert pred.is_contiguous() and pred.dtype == torch.bfloat16 and y.is_contiguous() and y.dtype == torch.float32
    acc = torch.zeros(1 + 4 * d_, device=pred.device, dtype=torch.float64)
    _call("a2v_d2v_loss_fused", pred, L.dtype_code(pred), _p(pred), _p(y), _p(mask_u8),
          _p(pred) if grad_coef is not None else None, C.c_int64(r), t, clones, d_, C.c_float(scale),
          C.c_float(grad_coef if grad_coef is not None else 0.0), C.c_void_p(acc.data_ptr()),
          C.c_void_p(acc.data_ptr() + 8))
    return acc[0:1], acc[1:].view(4, d_)


def d2v_loss_bwd(pred, y, mask_u8, clones, scale, grad_out: Optional[torch.Tensor]):
    r, t, d_ = pred.shape
    dpred = torch.empty_like(pred)
    if grad_out is not None:
        assert grad_out.dtype == torch.float32 and grad_out.numel() == 1
    _call("a2v_d2v_loss_bwd", pred, L.dtype_code(pred), _p(pred), _p(y), _p(mask_u8), _p(dpred), C.c_int64(r), t,
          clones, d_, C.c_float(scale), _p(grad_out))
    return dpred


# ----------------------------------------------------------------------------- utilities
def colsum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    c = x.shape[-1]
    assert x.is_contiguous() and out.dtype == torch.float32 and out.numel() == c
    _call("a2v_colsum", x, L.dtype_code(x), _p(x), _p(out), C.c_int64(x.numel() // c), c)
    return out


def dgelu_mul(dh: torch.Tensor, u: torch.Tensor, out: Optional[torch.Tensor] = None,
              colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = dh * GELU'(u) (in place on ``dh`` when ``out`` is None); ``colsum`` (fp32, one per column of the
    last dimension) optionally accumulates the column sums of the result in the same pass."""
    assert dh.is_contiguous() and u.is_contiguous() and dh.shape == u.shape and dh.dtype == u.dtype
    out = dh if out is None else out
    if colsum is not None:
        c = dh.shape[-1]
        assert colsum.dtype == torch.float32 and colsum.numel() == c
        _call("a2v_dgelu_mul_colsum", dh, L.dtype_code(dh), _p(dh), _p(u), _p(out), C.c_int64(dh.numel() // c), c,
              _p(colsum))
        return out
    _call("a2v_dgelu_mul", dh, L.dtype_code(dh), _p(dh), _p(u), _p(out), C.c_int64(dh.numel()))
    return out


def cast_strided(src: torch.Tensor, dims, strides, offset=0, out_dtype=torch.bfloat16) -> torch.Tensor:
    """Dense tensor of shape ``dims`` (<= 4-D) with out[i...] = src.flat[offset + sum i_d*strides[d]]."""
    dims = list(dims)
    strides = list(strides)
    while len(dims) < 4:
        dims.insert(0, 1)
        strides.insert(0, 0)
    out = torch.empty(dims, device=src.device, dtype=out_dtype)
    dd = (C.c_int64 * 4)(*dims)
    ss = (C.c_int64 * 4)(*strides)
    _call("a2v_cast_strided", src, L.dtype_code(src), L.dtype_code(out), _p(src), _p(out), dd, ss, C.c_int64(offset))
    return out


def relayout(src: torch.Tensor, dst: torch.Tensor, dims, in_strides, in_off, out_strides, out_off,
             accumulate: bool = False) -> torch.Tensor:
    """dst.flat[out_off + i.out_strides] (+)= src.flat[in_off + i.in_strides] over a <= 4-D index space."""
    dims, in_strides, out_strides = list(dims), list(in_strides), list(out_strides)
    while len(dims) < 4:
        dims.insert(0, 1)
        in_strides.insert(0, 0)
        out_strides.insert(0, 0)
    dd = (C.c_int64 * 4)(*dims)
    si = (C.c_int64 * 4)(*in_strides)
    so = (C.c_int64 * 4)(*out_strides)
    _call("a2v_relayout", src, L.dtype_code(src), L.dtype_code(dst), _p(src), _p(dst), dd, si, C.c_int64(in_off), so,
          C.c_int64(out_off), int(accumulate))
    return dst


class RelayoutItem(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p),
        ("out", C.c_void_p),
        ("dims", C.c_int64 * 4),
        ("in_strides", C.c_int64 * 4),
        ("out_strides", C.c_int64 * 4),
        ("in_offset", C.c_int64),
        ("out_offset", C.c_int64),
        ("in_dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("accumulate", C.c_int32),
        ("zero_src", C.c_int32),
    ]


class RelayoutTable:
    """A device-resident table of re-layout items (see a2v_relayout_batch): built once for a fixed set of
    persistent source / destination buffers, replayed with one launch. Keeps the tensors alive."""

    def __init__(self, device):
        self.device = device
        self.items: List[RelayoutItem] = []
        self.keep: list = []
        self.table: Optional[torch.Tensor] = None

    def add(self, src: torch.Tensor, dst: torch.Tensor, dims, in_strides, in_off, out_strides, out_off,
            accumulate: bool = False, zero_src: bool = False) -> None:
        dims, in_strides, out_strides = list(dims), list(in_strides), list(out_strides)
        while len(dims) < 4:
            dims.insert(0, 1)
            in_strides.insert(0, 0)
            out_strides.insert(0, 0)
        total = 1
        for v in dims:
            assert v > 0
            total *= v
        assert total < 2 ** 31, "relayout item too large for the batched kernel"
        it = RelayoutItem()
        it.in_, it.out = src.data_ptr(), dst.data_ptr()
        it.dims = (C.c_int64 * 4)(*dims)
        it.in_strides = (C.c_int64 * 4)(*in_strides)
        it.out_strides = (C.c_int64 * 4)(*out_strides)
        it.in_offset, it.out_offset = in_off, out_off
        it.in_dtype, it.out_dtype = L.dtype_code(src), L.dtype_code(dst)
        it.accumulate, it.zero_src = int(accumulate), int(zero_src)
        self.items.append(it)
        self.keep += [src, dst]
        self.table = None

    def run(self, blocks_per_item: int = 48) -> None:
        if not self.items:
            return
        if self.table is None:
            raw = b"".join(bytes(it) for it in self.items)
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        _call("a2v_relayout_batch", self.table, _p(self.table), len(self.items), blocks_per_item)


def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert src.dtype == torch.float32 and src.is_contiguous() and src.numel() % 4 == 0
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    _call("a2v_cast_f32_to_bf16", src, _p(src), _p(out), C.c_int64(src.numel()))
    return out


def cast_f32(src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    assert src.dtype == torch.bfloat16 and out.dtype == torch.float32 and src.numel() == out.numel() and src.numel() % 8 == 0
    _call("a2v_cast_bf16_to_f32", src, _p(src), _p(out), C.c_int64(src.numel()))
    return out


def split3(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 (rows, K) -> bf16 (rows, 3K) [pattern 0/1] or (3*rows, K) [pattern 2/3]."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    k = x.shape[-1]
    rows = x.numel() // k
    shape = (rows, 3 * k) if pattern < 2 else (3 * rows, k)
    out = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    _call("a2v_split3", x, _p(x), _p(out), C.c_int64(rows), k, pattern)
    return out


def ema_step(student: torch.Tensor, shadow: torch.Tensor, teacher_lp: Optional[torch.Tensor], decay: float) -> None:
    assert student.dtype == torch.float32 and shadow.dtype == torch.float32 and student.numel() == shadow.numel()
    _call("a2v_ema_step", student, _p(student), _p(shadow), _p(teacher_lp), C.c_int64(student.numel()),
          C.c_float(decay))


def adamw_step(p, g, m, v, p_lp, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, wd_mask=None) -> None:
    if wd_mask is not None:
        assert wd_mask.dtype == torch.uint8 and wd_mask.numel() * 4 == p.numel()
    _call("a2v_adamw_step", p, _p(p), _p(g), _p(m), _p(v), _p(p_lp), C.c_int64(p.numel()), C.c_float(lr),
          C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_float(weight_decay), int(step), _p(grad_scale),
          _p(wd_mask))


def sumsq(x: torch.Tensor, out: torch.Tensor) -> None:
    assert x.dtype == torch.float32 and out.dtype == torch.float64
    _call("a2v_sumsq", x, _p(x), C.c_int64(x.numel()), _p(out))


def clip_coef(sumsq_t: torch.Tensor, denom: Optional[torch.Tensor], numer: float, max_norm: float,
              out2: torch.Tensor) -> None:
    _call("a2v_clip_coef", sumsq_t, _p(sumsq_t), _p(denom), C.c_float(numer), C.c_float(max_norm), _p(out2))


# ----------------------------------------------------------------------------- sinc / im2col / mixup
def sinc_filters_fwd(low_hz, band_hz, n_, window_, k, min_low, min_band, sr):
    c = low_hz.numel()
    filt = torch.empty(128, k, device=low_hz.device, dtype=torch.float32)
    _call("a2v_sinc_filters_fwd", low_hz, _p(low_hz), _p(band_hz), _p(n_), _p(window_), c, k, C.c_float(min_low),
          C.c_float(min_band), C.c_float(sr), _p(filt))
    return filt


def sinc_filters_bwd(low_hz, band_hz, n_, window_, k, min_low, min_band, sr, dfilt, dlow, dband):
    c = low_hz.numel()
    _call("a2v_sinc_filters_bwd", low_hz, _p(low_hz), _p(band_hz), _p(n_), _p(window_), c, k, C.c_float(min_low),
          C.c_float(min_band), C.c_float(sr), _p(dfilt), _p(dlow), _p(dband))


def sinc_conv_fwd(x, filt, out_dtype):
    b, n = x.shape
    k = filt.shape[1]
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(b, n, 128, device=x.device, dtype=out_dtype)
    _call("a2v_sinc_conv_fwd", x, L.dtype_code(y), _p(x), _p(filt), _p(y), b, n, k)
    return y


def sinc_conv_wgrad(x, dy, k):
    b, n = x.shape
    dfilt = torch.zeros(128, k, device=x.device, dtype=torch.float32)
    _call("a2v_sinc_conv_wgrad", x, L.dtype_code(dy), _p(x), _p(dy), _p(dfilt), b, n, k)
    return dfilt


def im2col(x, k, stride, pad, t_out):
    b, t_in, c = x.shape
    assert x.is_contiguous()
    col = torch.empty(b, t_out, k * c, device=x.device, dtype=x.dtype)
    _call("a2v_im2col", x, L.dtype_code(x), _p(x), _p(col), b, t_in, t_out, c, k, stride, pad)
    return col


def col2im(dcol, k, stride, pad, t_in):
    b, t_out, kc = dcol.shape
    c = kc // k
    dx = torch.empty(b, t_in, c, device=dcol.device, dtype=dcol.dtype)
    _call("a2v_col2im", dcol, L.dtype_code(dcol), _p(dcol), _p(dx), b, t_in, t_out, c, k, stride, pad)
    return dx


def mixup_gain(x, hann, aweight, n_fft, hop, min_db=-80.0):
    b, n = x.shape
    w = (n - n_fft) // hop + 1
    g = torch.empty(b, w, device=x.device, dtype=torch.float32)
    _call("a2v_mixup_gain", x, _p(x), _p(hann), _p(aweight), b, n, n_fft, hop, C.c_float(min_db), _p(g))
    return g


def mixup_apply(x, perm_i32, gain_db, r):
    b, n = x.shape
    out = torch.empty_like(x)
    p = torch.empty(b, device=x.device, dtype=torch.float32)
    _call("a2v_mixup_apply", x, _p(x), _p(perm_i32), _p(gain_db), b, n, gain_db.shape[1], C.c_float(r), _p(out), _p(p))
    return out, p


# ----------------------------------------------------------------------------- finetune head / criterion
def layer_mean_head_fwd(layers: Sequence[torch.Tensor], w: torch.Tensor, bias: Optional[torch.Tensor],
                        save_mean: bool = True):
    """(logits (rows, C) fp32, xmean (rows, D) or None): mean over the given layer outputs, then Linear(D, C)."""
    k = len(layers)
    d_ = layers[0].shape[-1]
    rows = layers[0].numel() // d_
    for x in layers:
        assert x.is_contiguous() and x.shape == layers[0].shape and x.dtype == layers[0].dtype
    assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[1] == d_
    c = w.shape[0]
    dev = layers[0].device
    ptrs = h2d_async(torch.tensor([x.data_ptr() for x in layers], dtype=torch.int64), dev)
    xmean = torch.empty(rows, d_, device=dev, dtype=layers[0].dtype) if save_mean else None
    logits = torch.empty(rows, c, device=dev, dtype=torch.float32)
    _call("a2v_layer_mean_head_fwd", layers[0], L.dtype_code(layers[0]), _p(ptrs), k, C.c_int64(rows), d_, c, _p(w),
          _p(bias), _p(xmean), _p(logits))
    return logits, xmean


def head_bwd(dlogits: torch.Tensor, xmean: torch.Tensor, w: torch.Tensor, k: int, dw: torch.Tensor,
             db: Optional[torch.Tensor], want_g: bool = True) -> Optional[torch.Tensor]:
    rows, d_ = xmean.shape
    c = w.shape[0]
    assert dlogits.dtype == torch.float32 and dlogits.is_contiguous() and dlogits.numel() == rows * c
    assert dw.dtype == torch.float32 and dw.shape == w.shape
    g = torch.empty_like(xmean) if want_g else None
    _call("a2v_head_bwd", xmean, L.dtype_code(xmean), _p(dlogits), _p(xmean), _p(w), k, C.c_int64(rows), d_, c, _p(g),
          _p(dw), _p(db))
    return g


def focal_loss_fwd(logits, targets, *, perm=None, rows_per_clip=0, r=1.0, alpha=0.25, gamma=2.0, threshold=0.5,
                   want_unreduced=False, want_mixed_targets=False):
    """Returns (loss_sum double[1], counters int64[5] = tp/fp/tn/fn/n_correct, unreduced loss or None, mixed targets or
    None)."""
    rows, c = logits.shape
    assert logits.dtype == torch.float32 and targets.dtype == torch.float32 and logits.is_contiguous()
    assert targets.is_contiguous() and targets.numel() == logits.numel()
    dev = logits.device
    loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
    counters = torch.zeros(5, device=dev, dtype=torch.int64)
    un = torch.empty_like(logits) if want_unreduced else None
    mt = torch.empty_like(logits) if want_mixed_targets else None
    _call("a2v_focal_loss_fwd", logits, _p(logits), _p(targets), _p(perm), C.c_int64(rows), c, int(rows_per_clip),
          C.c_float(r), C.c_float(alpha), C.c_float(gamma), C.c_float(threshold), _p(loss_sum), _p(un), _p(mt),
          _p(counters))
    return loss_sum, counters, un, mt


def focal_loss_bwd(logits, targets, *, perm=None, rows_per_clip=0, r=1.0, alpha=0.25, gamma=2.0, grad_out=None):
    rows, c = logits.shape
    dlogits = torch.empty_like(logits)
    _call("a2v_focal_loss_bwd", logits, _p(logits), _p(targets), _p(perm), C.c_int64(rows), c, int(rows_per_clip),
          C.c_float(r), C.c_float(alpha), C.c_float(gamma), _p(grad_out), _p(dlogits))
    return dlogits


def channel_mask_(x: torch.Tensor, chmask_u8: torch.Tensor, rows_per_clip: int) -> torch.Tensor:
    d_ = x.shape[-1]
    assert x.is_contiguous() and chmask_u8.dtype == torch.uint8 and chmask_u8.shape[-1] == d_
    _call("a2v_channel_mask", x, L.dtype_code(x), _p(x), _p(chmask_u8), C.c_int64(x.numel() // d_), int(rows_per_clip), d_)
    return x


# ----------------------------------------------------------------------------- input pipeline
def clip_layer_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Per-clip F.layer_norm(feats, feats.shape) of a collated (B, N) fp32 batch."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    y = torch.empty_like(x)
    _call("a2v_clip_layer_norm", x, _p(x), _p(y), x.shape[0], x.shape[1], C.c_float(eps))
    return y


def frame_labels(offsets, start, end, cat, foc, batch: int, frames: int, classes: int, wav_len: int, focal_class: int):
    """(B, T, classes) fp32 multi-hot frame labels from flat int32 interval arrays (see a2v_frame_labels)."""
    out = torch.empty(batch, frames, classes, device=offsets.device, dtype=torch.float32)
    _call("a2v_frame_labels", offsets, _p(offsets), _p(start), _p(end), _p(cat), _p(foc), batch, frames, classes, wav_len,
          focal_class, _p(out))
    return out
