"""Parameter inventory, flat storage and GEMM weight packing for the pretraining path.

State-dict key names and shapes are the reference's checkpoint ABI (module
registration order of nn/modalities/base.py:92-135, nn/modalities/audio.py:71-149 and
nn/data2vec2.py:236-277). All student parameters live in ONE flat fp32 buffer (with one flat fp32
gradient buffer of the same layout), the parameters the EMA teacher shares come first so that the
teacher update (fairseq EMAModule.step as called from data2vec2.py:408) is a single fused launch
over a contiguous range, and gradient buckets for the NCCL all-reduce are contiguous slices.

"Packs" are the bf16 operand layouts the tcgen05 GEMM reads: tap-major conv weights, group-padded
decoder weights (48 real channels in 64-wide groups), transposed copies for the data gradients.
They are rebuilt from the fp32 masters by a2v_relayout after every optimizer step.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .config import Data2VecMultiConfig, parse_conv_layers

ENC = "modality_encoders.AUDIO."
ALIGN = 8  # elements; keeps every tensor 32-byte (fp32) / 16-byte (bf16 copy) aligned


def _block_shapes(prefix: str, d: int, hidden: int) -> Dict[str, Tuple[int, ...]]:
    return {
        prefix + "norm1.weight": (d,), prefix + "norm1.bias": (d,),
        prefix + "attn.qkv.weight": (3 * d, d), prefix + "attn.qkv.bias": (3 * d,),
        prefix + "attn.proj.weight": (d, d), prefix + "attn.proj.bias": (d,),
        prefix + "norm2.weight": (d,), prefix + "norm2.bias": (d,),
        prefix + "mlp.fc1.weight": (hidden, d), prefix + "mlp.fc1.bias": (hidden,),
        prefix + "mlp.fc2.weight": (d, hidden), prefix + "mlp.fc2.bias": (d,),
    }


def pos_kernel(cfg: Data2VecMultiConfig) -> int:
    a = cfg.modalities.audio
    return max(3, a.conv_pos_width // a.conv_pos_depth)  # nn/modalities/audio.py:91


def student_param_shapes(cfg: Data2VecMultiConfig) -> Dict[str, Tuple[int, ...]]:
    """name -> shape in the reference's named_parameters() order."""
    a = cfg.modalities.audio
    d, hidden = cfg.embed_dim, int(cfg.embed_dim * cfg.mlp_ratio)
    layers = parse_conv_layers(a.conv_feature_layers)
    s: Dict[str, Tuple[int, ...]] = {}
    # nn/modalities/base.py:116-134 (learned_alibi_scale_per_layer is refused by the engine)
    s[ENC + "alibi_scale"] = (1, 1, a.num_alibi_heads if a.learned_alibi_scale_per_head else 1, 1, 1)
    c0, _k0, _ = layers[0]
    le = ENC + "local_encoder.conv_layers."
    s[le + "0.0.low_hz_"] = (c0, 1)
    s[le + "0.0.band_hz_"] = (c0, 1)
    s[le + "0.2.1.weight"] = (c0,)
    s[le + "0.2.1.bias"] = (c0,)
    s[le + "0.3.p_swish_alpha"] = (1, c0, 1)
    s[le + "0.3.p_swish_beta"] = (1, c0, 1)
    cin = c0
    for i, (c, k, _st) in enumerate(layers[1:], start=1):
        s[le + f"{i}.0.weight"] = (c, cin, k)
        s[le + f"{i}.2.1.weight"] = (c,)
        s[le + f"{i}.2.1.bias"] = (c,)
        cin = c
    s[ENC + "project_features.1.weight"] = (cin,)
    s[ENC + "project_features.1.bias"] = (cin,)
    s[ENC + "project_features.2.weight"] = (d, cin)
    s[ENC + "project_features.2.bias"] = (d,)
    kp = pos_kernel(cfg)
    for i in range(1, a.conv_pos_depth + 1):
        s[ENC + f"relative_positional_encoder.{i}.0.weight"] = (d, d // a.conv_pos_groups, kp)
        s[ENC + f"relative_positional_encoder.{i}.0.bias"] = (d,)
    for j in range(a.prenet_depth):
        s.update(_block_shapes(ENC + f"context_encoder.blocks.{j}.", d, hidden))
    s[ENC + "context_encoder.norm.weight"] = (d,)
    s[ENC + "context_encoder.norm.bias"] = (d,)
    dec = a.decoder
    for l in range(dec.decoder_layers):
        cin_l = d if l == 0 else dec.decoder_dim
        s[ENC + f"decoder.blocks.{l}.0.weight"] = (dec.decoder_dim, cin_l // dec.decoder_groups, dec.decoder_kernel)
        s[ENC + f"decoder.blocks.{l}.0.bias"] = (dec.decoder_dim,)
    s[ENC + "decoder.proj.weight"] = (d, dec.decoder_dim)
    s[ENC + "decoder.proj.bias"] = (d,)
    for j in range(cfg.depth):
        s.update(_block_shapes(f"blocks.{j}.", d, hidden))
    return s


def is_teacher_key(k: str) -> bool:
    """The EMA teacher has no local_encoder / project_features / decoder (nn/data2vec2.py:377-381)."""
    return not (k.startswith(ENC + "local_encoder.") or k.startswith(ENC + "project_features.")
                or k.startswith(ENC + "decoder."))


def no_decay(k: str, shape: Tuple[int, ...]) -> bool:
    """nn/data2vec2.py:318-322: weight_decay_scale 0 for 1-D tensors, biases, alibi_scale and p_swish."""
    return len(shape) == 1 or k.endswith(".bias") or "alibi_scale" in k or "p_swish" in k


class FlatParams:
    """One flat fp32 buffer (+ gradient buffer) with named views."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, *, with_grad: bool, order: Optional[List[str]] = None):
        self.shapes = dict(shapes)
        self.names = list(order) if order is not None else list(shapes.keys())
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.names:
            self.offsets[n] = off
            numel = int(np.prod(shapes[n]))
            off += (numel + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.data = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32) if with_grad else None

    def numel(self, n: str) -> int:
        return int(np.prod(self.shapes[n]))

    def view(self, n: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = self.data if buf is None else buf
        o = self.offsets[n]
        return buf[o:o + self.numel(n)].view(self.shapes[n])

    def gview(self, n: str) -> torch.Tensor:
        return self.view(n, self.grad)

    def range_of(self, names: List[str]) -> Tuple[int, int]:
        lo = min(self.offsets[n] for n in names)
        hi = max(self.offsets[n] + (self.numel(n) + ALIGN - 1) // ALIGN * ALIGN for n in names)
        return lo, hi


def student_layout(cfg: Data2VecMultiConfig) -> Tuple[Dict[str, Tuple[int, ...]], List[str], List[str]]:
    """(shapes, flat order, teacher-shared names). Shared parameters first, in reference order."""
    shapes = student_param_shapes(cfg)
    shared = [k for k in shapes if is_teacher_key(k)]
    rest = [k for k in shapes if not is_teacher_key(k)]
    return shapes, shared + rest, shared


# --------------------------------------------------------------------------------------------
# weight packs
# --------------------------------------------------------------------------------------------
class Pack:
    """A GEMM-operand view of one parameter: 4-D index map from the reference layout into a dense,
    possibly padded buffer. ``split_k`` is the K granularity of the fp32-mode hi/lo split."""

    __slots__ = ("src", "dims", "in_strides", "in_off", "out_strides", "out_shape", "split_k", "buf", "fbuf", "is_bias",
                 "gt_strides", "gt_shape", "extra_parts")

    def __init__(self, src, dims, in_strides, in_off, out_strides, out_shape, split_k, is_bias=False):
        self.src, self.dims, self.in_strides, self.in_off = src, list(dims), list(in_strides), in_off
        self.out_strides, self.out_shape, self.split_k = list(out_strides), tuple(out_shape), split_k
        self.buf = None
        self.fbuf = None
        self.is_bias = is_bias
        self.gt_strides = None  # strides of the same 4-D index space in the TRANSPOSED conv weight-gradient buffer
        self.gt_shape = None
        self.extra_parts = []   # further (dims, in_strides, in_off, out_strides, out_off) maps into the same buffer

    @property
    def parts(self):
        """Every 4-D index map that fills this pack's buffer: (dims, in_strides, in_off, out_strides, out_off)."""
        return [(self.dims, self.in_strides, self.in_off, self.out_strides, 0)] + list(self.extra_parts)


def pack_linear_t(name, n_out, k_in) -> Pack:
    """W (n_out, k_in) -> W^T (k_in, n_out): B operand of the data-gradient GEMM."""
    return Pack(name, (1, 1, k_in, n_out), (0, 0, 1, k_in), 0, (0, 0, n_out, 1), (k_in, n_out), n_out)


def pack_conv_fwd(name, groups, ng, cg, k, ngp=None, cgp=None) -> Pack:
    """torch Conv1d weight (G*ng, cg, k) -> (G*ngp, k*cgp), tap-major, channel-minor, zero padded."""
    ngp, cgp = ngp or ng, cgp or cg
    pk = Pack(name, (groups, ng, k, cg), (ng * cg * k, cg * k, 1, k), 0, (ngp * k * cgp, k * cgp, cgp, 1),
              (groups * ngp, k * cgp), cgp)
    # gemm.conv_wgrad_tn writes (G*k*cgp, ngp): element (g, n, j, c) at ((g*k + j)*cgp + c)*ngp + n
    pk.gt_strides = [k * cgp * ngp, 1, cgp * ngp, ngp]
    pk.gt_shape = (groups * k * cgp, ngp)
    return pk


def pack_conv_dgrad(name, groups, ng, cg, k, ngp=None, cgp=None) -> Pack:
    """Flipped + transposed weight for the input gradient of a stride-1 conv expressed as a conv of dy:
    wd[g*cgp + c, j'*ngp + n] = W[g*ng + n, c, k-1-j']."""
    ngp, cgp = ngp or ng, cgp or cg
    return Pack(name, (groups, cg, k, ng), (ng * cg * k, k, -1, cg * k), k - 1, (cgp * k * ngp, k * ngp, ngp, 1),
                (groups * cgp, k * ngp), ngp)


def pack_col_t(name, c_out, cin, k, cinp) -> Pack:
    """(c_out, cin, k) -> (k*cinp, c_out): B operand of dcol = dy @ Wcol for the im2col convs."""
    return Pack(name, (1, k, cin, c_out), (0, 1, k, cin * k), 0, (0, cinp * c_out, c_out, 1), (k * cinp, c_out), c_out)


def pack_conv_dgrad_strided(name, c_out, cin, k, stride, pad, cinp) -> Pack:
    """B operand of gemm.strided_conv_dgrad: torch Conv1d weight (c_out, cin, k) -> (stride*cinp, tpb*c_out) with
    wd[r*cinp + c, u*c_out + n] = W[n, c, (r + pad) % stride + stride*u]  (zero where that tap index is >= k);
    tpb = ceil(k / stride) taps per output column block r. One 4-D index map per block r."""
    tpb = -(-k // stride)
    parts = []
    for r in range(stride):
        j0 = (r + pad) % stride
        nu = len(range(j0, k, stride))
        if nu == 0:
            continue
        parts.append(((1, cin, nu, c_out), (0, k, stride, cin * k), j0, (0, tpb * c_out, c_out, 1), r * cinp * tpb * c_out))
    d0 = parts[0]
    pk = Pack(name, d0[0], d0[1], d0[2], d0[3], (stride * cinp, tpb * c_out), c_out)
    assert d0[4] == 0 or True
    # the first part may not start at out offset 0 (block 0 always has taps, so it does)
    pk.extra_parts = parts[1:]
    if d0[4] != 0:
        raise ValueError("pack_conv_dgrad_strided: block 0 has no tap")
    return pk


def pack_cols_padded(name, n_out, groups, cg, cgp) -> Pack:
    """Linear weight (n_out, G*cg) whose INPUT is group padded -> (n_out, G*cgp)."""
    return Pack(name, (1, n_out, groups, cg), (0, groups * cg, cg, 1), 0, (0, groups * cgp, cgp, 1),
                (n_out, groups * cgp), groups * cgp)


def pack_cols_padded_t(name, n_out, groups, cg, cgp) -> Pack:
    """Transpose of :func:`pack_cols_padded`: (G*cgp, n_out)."""
    return Pack(name, (1, groups, cg, n_out), (0, cg, 1, groups * cg), 0, (0, cgp * n_out, n_out, 1),
                (groups * cgp, n_out), n_out)


def pack_bias_padded(name, groups, ng, ngp) -> Pack:
    return Pack(name, (1, 1, groups, ng), (0, 0, ng, 1), 0, (0, 0, ngp, 1), (groups * ngp,), 0, is_bias=True)


def materialize(p: Pack, src: torch.Tensor, fp32_mode: bool) -> torch.Tensor:
    """(Re)build the operand buffer of ``p`` from the fp32 master tensor ``src``."""
    dev = src.device
    if p.is_bias:
        if p.buf is None:
            p.buf = torch.zeros(p.out_shape, device=dev, dtype=torch.float32)
        ops.relayout(src, p.buf, p.dims, p.in_strides, p.in_off, p.out_strides, 0)
        return p.buf
    if not fp32_mode:
        if p.buf is None:
            p.buf = torch.zeros(p.out_shape, device=dev, dtype=torch.bfloat16)
        for dims, ist, ioff, ost, ooff in p.parts:
            ops.relayout(src, p.buf, dims, ist, ioff, ost, ooff)
        return p.buf
    if p.fbuf is None:
        p.fbuf = torch.zeros(p.out_shape, device=dev, dtype=torch.float32)
    for dims, ist, ioff, ost, ooff in p.parts:
        ops.relayout(src, p.fbuf, dims, ist, ioff, ost, ooff)
    p.buf = ops.split3(p.fbuf.view(-1, p.split_k), 1).view(p.out_shape[0], 3 * p.out_shape[1])
    return p.buf


def unpack_grad(p: Pack, packed_grad: torch.Tensor, dst: torch.Tensor, transposed: bool = False) -> None:
    """Inverse index map: ADD the packed-layout fp32 gradient into the reference-layout view (the packed
    buffers are transient per backward; the flat gradient buffer is the only accumulator)."""
    strides = p.gt_strides if transposed else p.out_strides
    ops.relayout(packed_grad, dst, p.dims, strides, 0, p.in_strides, p.in_off, accumulate=True)


# --------------------------------------------------------------------------------------------
# initialisation (module-level defaults of the reference; parity tests load explicit weights)
# --------------------------------------------------------------------------------------------
def sinc_mel_init(c0: int, k0: int, sample_rate: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn/sinc.py:225-253: mel-spaced low cut-offs and band widths."""
    min_low_hz = 50.0
    min_band_hz = float(math.ceil(sample_rate / k0))
    high_hz = sample_rate / 2 - (min_low_hz + min_band_hz)
    mel = torch.linspace(2595 * np.log10(1 + min_low_hz / 700), 2595 * np.log10(1 + high_hz / 700), c0 + 1)
    hz = 700 * (10 ** (mel / 2595) - 1)
    return hz[:-1].unsqueeze(1).float(), (hz[1:] - hz[:-1]).unsqueeze(1).float()


def sinc_buffers(k0: int, sample_rate: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn/sinc.py:264-276: half Hamming window on linspace(0, k/2-1, int(k/2)) and the time axis
    2*pi*arange(-(k-1)/2, 0)/sr."""
    n_lin = torch.linspace(0, (k0 / 2) - 1, steps=int(k0 / 2))
    window = 0.53836 - 0.46164 * torch.cos(2 * math.pi * n_lin / k0)
    n = (k0 - 1) / 2.0
    n_ = 2 * math.pi * torch.arange(-n, 0) / sample_rate
    return n_.float().contiguous(), window.float().contiguous()


def default_init(cfg: Data2VecMultiConfig, seed: int) -> Dict[str, torch.Tensor]:
    """Reference-style random init on the host: Linear weights N(0, 0.02) with zero bias
    (fairseq init_bert_params, applied at nn/data2vec2.py:286-288), kaiming-normal feature-extractor
    convs (nn/utils.py:1085-1092), N(0, sqrt(4/(k*D))) positional convs with zero bias
    (fairseq-style conv_pos init), LayerNorm weight 1 / bias 0, PSwish alpha 2 / beta 0
    (nn/utils.py:1413-1428), mel-spaced sinc parameters, alibi_scale = cfg value."""
    g = torch.Generator().manual_seed(seed)
    a = cfg.modalities.audio
    layers = parse_conv_layers(a.conv_feature_layers)
    low, band = sinc_mel_init(layers[0][0], layers[0][1], a.sample_rate)
    out: Dict[str, torch.Tensor] = {}
    for k, shape in student_param_shapes(cfg).items():
        if k.endswith("low_hz_"):
            t = low.clone()
        elif k.endswith("band_hz_"):
            t = band.clone()
        elif k.endswith("p_swish_alpha"):
            t = torch.full(shape, 2.0)
        elif k.endswith("p_swish_beta"):
            t = torch.zeros(shape)
        elif k.endswith("alibi_scale"):
            t = torch.full(shape, float(a.alibi_scale))
        elif len(shape) == 1:
            t = torch.ones(shape) if k.endswith("weight") else torch.zeros(shape)
        elif "local_encoder" in k:
            fan_in = shape[1] * shape[2]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif "relative_positional_encoder" in k or "decoder.blocks" in k:
            fan_in = shape[1] * shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = torch.randn(shape, generator=g) * 0.02
        out[k] = t.float()
    return out
