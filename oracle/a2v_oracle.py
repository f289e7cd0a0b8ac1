"""TEST INFRASTRUCTURE ONLY (the oracle). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product (animal2vec_b200/)
never does.

CPU fp32 restatement, in plain functional torch, of the reference's data2vec2-style
pretraining step (animal2vec, /root/reference). Every function cites the reference lines it
follows. It is pinned against the *real* reference code: tests/golden/make_golden.py imports
/root/reference/nn under oracle/ref_shims.py, runs the reference model on seeded inputs and
stores the outputs as fixtures; tests/test_oracle.py checks this file against those fixtures.
The third-party pieces the reference calls (fairseq compute_mask_indices / EMAModule, timm
Mlp) are absent from /root/reference; they are restated from their published behaviour
 -- for those the parity is "unpinned upstream" and anchored on the
reference's call sites.

Parameters are a flat ``dict[str, Tensor]`` keyed exactly like the reference's
``state_dict()`` (checkpoint ABI).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SHIPPED_CONV_LAYERS = "[(127, 63, 1)] +[(512, 10, 5)] + [(512, 3, 2)] * 3 + [(512, 3, 1)] + [(512, 2, 1)] * 2"


@dataclass
class OracleConfig:
    """The subset of Data2VecMultiConfig / D2vAudioConfig / D2vDecoderConfig the pretraining path
    reads (reference nn/data2vec2.py:56-166, nn/modalities/base.py:30-72, audio.py:29-51,
    modules.py:34-47); defaults = the shipped large recipe."""

    embed_dim: int = 1024
    num_heads: int = 16
    depth: int = 16
    prenet_depth: int = 8
    mlp_ratio: float = 4.0
    norm_eps: float = 1e-5
    clone_batch: int = 12
    average_top_k_layers: int = 16
    ema_decay: float = 0.9997
    ema_end_decay: float = 1.0
    ema_anneal_end_step: int = 300000
    seed: int = 1
    sample_rate: int = 8000
    conv_feature_layers: str = SHIPPED_CONV_LAYERS
    conv_pos_width: int = 95
    conv_pos_depth: int = 5
    conv_pos_groups: int = 16
    mask_prob: float = 1.5
    mask_length: int = 2
    mask_noise_std: float = 0.01
    decoder_dim: int = 768
    decoder_groups: int = 16
    decoder_kernel: int = 7
    decoder_layers: int = 4
    decoder_input_dropout: float = 0.1
    encoder_dropout: float = 0.1
    attention_dropout: float = 0.1
    post_mlp_drop: float = 0.1
    prenet_dropout: float = 0.1
    source_mixup: float = 0.5
    mixing_window_length: float = 0.05
    loss_scale: Optional[float] = None

    @property
    def conv_layers(self) -> List[Tuple[int, int, int]]:
        return eval(self.conv_feature_layers)  # the reference evals this string too (audio.py:68)

    @property
    def pos_kernel(self) -> int:
        return max(3, self.conv_pos_width // self.conv_pos_depth)  # audio.py:91


def large_config(**kw) -> OracleConfig:
    return OracleConfig(**kw)


def base_config(**kw) -> OracleConfig:
    """"animal2vec-base" as defined in SURVEY.md: dataclass defaults for the transformer
    (depth 8, dim 768, 12 heads, prenet 4) + everything else from the large YAML, top-K 8."""
    d = dict(embed_dim=768, num_heads=12, depth=8, prenet_depth=4, average_top_k_layers=8)
    d.update(kw)
    return OracleConfig(**d)


def tiny_config(**kw) -> OracleConfig:
    """Small configuration for golden fixtures: every code path of the large recipe, 64-wide
    feature extractor, 48-wide decoder groups (exercises the channel-group padding)."""
    d = dict(embed_dim=128, num_heads=2, depth=2, prenet_depth=2, clone_batch=3, average_top_k_layers=2,
             conv_feature_layers="[(127, 63, 1)] +[(64, 10, 5)] + [(64, 3, 2)] * 3 + [(64, 3, 1)] + [(64, 2, 1)] * 2",
             conv_pos_depth=2, conv_pos_width=38, conv_pos_groups=2, decoder_dim=96, decoder_groups=2,
             decoder_kernel=7, decoder_layers=2, ema_anneal_end_step=1000)
    d.update(kw)
    return OracleConfig(**d)


# ------------------------------------------------------------------------------------------------
# parameter inventory (state-dict keys / shapes) and a seed-deterministic init shared by the
# reference run (make_golden.py), the oracle and the CUDA path
# ------------------------------------------------------------------------------------------------
ENC = "modality_encoders.AUDIO."


def _block_shapes(prefix: str, d: int, hidden: int) -> Dict[str, Tuple[int, ...]]:
    return {
        prefix + "norm1.weight": (d,), prefix + "norm1.bias": (d,),
        prefix + "attn.qkv.weight": (3 * d, d), prefix + "attn.qkv.bias": (3 * d,),
        prefix + "attn.proj.weight": (d, d), prefix + "attn.proj.bias": (d,),
        prefix + "norm2.weight": (d,), prefix + "norm2.bias": (d,),
        prefix + "mlp.fc1.weight": (hidden, d), prefix + "mlp.fc1.bias": (hidden,),
        prefix + "mlp.fc2.weight": (d, hidden), prefix + "mlp.fc2.bias": (d,),
    }


def student_param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Keys in the order of the reference's named_parameters() (module registration order:
    modality_encoders.AUDIO.{local_encoder, project_features, relative_positional_encoder,
    context_encoder, decoder, alibi_scale}, then blocks) -- nn/modalities/base.py:92-135,
    audio.py:71-149, data2vec2.py:236-277."""
    d, hidden = cfg.embed_dim, int(cfg.embed_dim * cfg.mlp_ratio)
    s: Dict[str, Tuple[int, ...]] = {}
    s[ENC + "alibi_scale"] = (1, 1, cfg.num_heads, 1, 1)  # nn.Parameter set in the base ctor, before submodules? see note
    layers = cfg.conv_layers
    c0, k0, _ = layers[0]
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
    kp = cfg.pos_kernel
    for i in range(1, cfg.conv_pos_depth + 1):
        s[ENC + f"relative_positional_encoder.{i}.0.weight"] = (d, d // cfg.conv_pos_groups, kp)
        s[ENC + f"relative_positional_encoder.{i}.0.bias"] = (d,)
    for j in range(cfg.prenet_depth):
        s.update(_block_shapes(ENC + f"context_encoder.blocks.{j}.", d, hidden))
    s[ENC + "context_encoder.norm.weight"] = (d,)
    s[ENC + "context_encoder.norm.bias"] = (d,)
    dd = cfg.decoder_dim
    for l in range(cfg.decoder_layers):
        cin_l = d if l == 0 else dd
        s[ENC + f"decoder.blocks.{l}.0.weight"] = (dd, cin_l // cfg.decoder_groups, cfg.decoder_kernel)
        s[ENC + f"decoder.blocks.{l}.0.bias"] = (dd,)
    s[ENC + "decoder.proj.weight"] = (d, dd)
    s[ENC + "decoder.proj.bias"] = (d,)
    for j in range(cfg.depth):
        s.update(_block_shapes(f"blocks.{j}.", d, hidden))
    return s


def is_teacher_key(k: str) -> bool:
    """The EMA teacher drops local_encoder, project_features and decoder (data2vec2.py:377-381)."""
    return not (k.startswith(ENC + "local_encoder.") or k.startswith(ENC + "project_features.")
                or k.startswith(ENC + "decoder."))


def init_params(cfg: OracleConfig, seed: int) -> Dict[str, Tensor]:
    """Seed-deterministic random init with reference-like scales (NOT the reference's RNG
    stream): Linear/conv weights ~ N(0, sigma), norms near identity with jitter so that affine
    paths are exercised, sinc parameters mel-spaced exactly as nn/sinc.py:225-253."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    c0, k0, _ = cfg.conv_layers[0]
    min_band = float(np.ceil(cfg.sample_rate / k0).astype(int))
    high_hz = cfg.sample_rate / 2 - (50 + min_band)
    mel = torch.linspace(2595 * np.log10(1 + 50 / 700), 2595 * np.log10(1 + high_hz / 700), c0 + 1)
    hz = 700 * (10 ** (mel / 2595) - 1)
    for k, shape in student_param_shapes(cfg).items():
        if k.endswith("low_hz_"):
            t = hz[:-1].unsqueeze(1).clone()
        elif k.endswith("band_hz_"):
            t = (hz[1:] - hz[:-1]).unsqueeze(1).clone()
        elif k.endswith("p_swish_alpha"):
            t = 2.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("p_swish_beta"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("alibi_scale"):
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif ".norm" in k or ".2.1." in k or "project_features.1." in k:
            t = (1.0 if k.endswith("weight") else 0.0) + 0.05 * torch.randn(shape, generator=g)
        elif k.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        out[k] = t.float()
    return out


# ------------------------------------------------------------------------------------------------
# feature extractor (reference nn/sinc.py, nn/utils.py:1043-1163)
# ------------------------------------------------------------------------------------------------
def sinc_filters(low_hz_: Tensor, band_hz_: Tensor, kernel_size: int, sample_rate: int) -> Tensor:
    """nn/sinc.py:181-223 with the buffers of _init_sinc_conv (:264-276)."""
    min_low_hz = 50
    min_band_hz = float(np.ceil(sample_rate / kernel_size).astype(int))  # :79
    n_lin = torch.linspace(0, (kernel_size / 2) - 1, steps=int(kernel_size / 2))
    window_ = 0.53836 - 0.46164 * torch.cos(2 * math.pi * n_lin / kernel_size)
    n = (kernel_size - 1) / 2.0
    n_ = 2 * math.pi * torch.arange(-n, 0).view(1, -1) / sample_rate
    low = min_low_hz + torch.abs(low_hz_)
    high = torch.clamp(low + min_band_hz + torch.abs(band_hz_), min_low_hz, sample_rate / 2)
    band = (high - low)[:, 0]
    f_low = torch.matmul(low.float(), n_.float())
    f_high = torch.matmul(high.float(), n_.float())
    left = (torch.sin(f_high) - torch.sin(f_low)) / n_.float() * 2 * window_.float()
    center = 2 * band.view(-1, 1)
    band_pass = torch.cat([left, center, torch.flip(left, dims=[1])], dim=1)
    return band_pass / (2 * band[:, None])


def pswish(x: Tensor, alpha: Tensor, beta: Tensor) -> Tensor:
    return x * alpha * torch.sigmoid(beta * x)  # nn/utils.py:1430-1431


def feature_extractor(p: Dict[str, Tensor], cfg: OracleConfig, source: Tensor, taps: Optional[dict] = None) -> Tensor:
    """ConvFeatureExtractionModel.forward (nn/utils.py:1155-1163) for sinc_input + layer_norm +
    PSwish: returns (B, C, T). `taps` collects per-layer outputs."""
    le = ENC + "local_encoder.conv_layers."
    layers = cfg.conv_layers
    c0, k0, _ = layers[0]
    x = source.unsqueeze(1)
    filt = sinc_filters(p[le + "0.0.low_hz_"], p[le + "0.0.band_hz_"], k0, cfg.sample_rate)
    pad = (k0 - 1) // 2  # nn/sinc.py:316-337 with L_in = in_channels = 1 -> (k-1)/2 each side
    x = F.conv1d(F.pad(x, (pad, pad), mode="reflect"), filt.view(c0, 1, k0))
    x = F.layer_norm(x.transpose(1, 2), (c0,), p[le + "0.2.1.weight"], p[le + "0.2.1.bias"], 1e-5).transpose(1, 2)
    x = pswish(x, p[le + "0.3.p_swish_alpha"], p[le + "0.3.p_swish_beta"])
    if taps is not None:
        taps["fe_layer0"] = x
    for i, (c, k, st) in enumerate(layers[1:], start=1):
        w = p[le + f"{i}.0.weight"]
        if st == 1:
            total = k - 1  # padding="same": left = total // 2, the extra one goes right
            x = F.conv1d(F.pad(x, (total // 2, total - total // 2)), w)
        else:
            x = F.conv1d(x, w, stride=st, padding=int(np.ceil(st / 2)))  # nn/utils.py:1089
        x = F.layer_norm(x.transpose(1, 2), (c,), p[le + f"{i}.2.1.weight"], p[le + f"{i}.2.1.bias"], 1e-5)
        x = F.gelu(x.transpose(1, 2))
        if taps is not None:
            taps[f"fe_layer{i}"] = x
    return x


def local_features(p: Dict[str, Tensor], cfg: OracleConfig, source: Tensor, taps: Optional[dict] = None,
                   frozen_extractor: bool = False) -> Tensor:
    """ModalitySpecificEncoder.local_features (base.py:194-213) + project_features (audio.py:83-88).
    ``frozen_extractor``: local_grad_mult = 0 -- ONLY the conv extractor runs under no_grad (base.py:205-207);
    project_features (LayerNorm + Linear) keeps its gradient."""
    if frozen_extractor:
        with torch.no_grad():
            x = feature_extractor(p, cfg, source, taps).transpose(1, 2)
    else:
        x = feature_extractor(p, cfg, source, taps).transpose(1, 2)
    c = x.shape[-1]
    x = F.layer_norm(x, (c,), p[ENC + "project_features.1.weight"], p[ENC + "project_features.1.bias"], 1e-5)
    return F.linear(x, p[ENC + "project_features.2.weight"], p[ENC + "project_features.2.bias"])


# ------------------------------------------------------------------------------------------------
# masking (fairseq compute_mask_indices as called from base.py:370-425) -- integer work, bit exact
# ------------------------------------------------------------------------------------------------
def clone_ids(seed: int, ids: Tensor, clone_batch: int) -> Tensor:
    """base.py:246-259."""
    clone_hash = [int(hash((seed, ind)) % 1e10) for ind in range(clone_batch - 1)]
    clone_hash = torch.tensor([0] + clone_hash).long().view(1, -1)
    i = ids.repeat_interleave(clone_batch, 0)
    return (i.view(-1, clone_batch) + clone_hash.to(i)).view(-1)


def compute_mask_indices(bsz: int, all_sz: int, mask_prob: float, mask_length: int, seed: int, epoch: int,
                         indices: Tensor, min_masks: int = 1) -> np.ndarray:
    """fairseq.data.data_utils.compute_mask_indices, static lengths, overlap allowed,
    require_same_masks=True, mask_dropout=0."""
    mask = np.full((bsz, all_sz), False)
    idcs = []
    rng = None
    for i in range(bsz):
        rng = np.random.default_rng(int(hash((seed, epoch, indices[i].item())) % 1e6))
        sz = all_sz
        num_mask = max(min_masks, int(mask_prob * sz / float(mask_length) + rng.random()))
        min_len = mask_length
        if sz - min_len <= num_mask:
            min_len = sz - num_mask - 1
        starts = rng.choice(sz - min_len, num_mask, replace=False)
        idc = np.asarray([starts[j] + o for j in range(len(starts)) for o in range(mask_length)])
        idc = np.unique(idc[idc < sz])
        if len(idc) >= sz:
            raise ValueError("the entire sequence is masked")
        idcs.append(idc)
    target_len = min(len(m) for m in idcs)
    for i, idc in enumerate(idcs):
        if len(idc) > target_len:
            idc = rng.choice(idc, target_len, replace=False)  # NB: the LAST row's rng, as upstream
        mask[i, idc] = True
    return mask


def pretrain_mask(cfg: OracleConfig, b: int, t: int, ids: Tensor, num_updates: int) -> np.ndarray:
    """(B*clone_batch, T) bool mask of one forward (data2vec2.py:618-620, base.py:241-271,401-413)."""
    ids_c = clone_ids(cfg.seed, ids, cfg.clone_batch) if cfg.clone_batch > 1 else ids
    return compute_mask_indices(b * cfg.clone_batch, t, cfg.mask_prob, cfg.mask_length, cfg.seed, num_updates, ids_c)


# ------------------------------------------------------------------------------------------------
# transformer pieces (nn/modalities/modules.py, base.py ALiBi)
# ------------------------------------------------------------------------------------------------
def alibi_slopes(n: int) -> List[float]:
    """base.py:559-576."""
    def pow2(n):
        start = 2 ** (-(2 ** -(math.log2(n) - 3)))
        return [start * start ** i for i in range(n)]

    if math.log2(n).is_integer():
        return pow2(n)
    c = 2 ** math.floor(math.log2(n))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n - c]


def alibi_bias(cfg: OracleConfig, scale: Tensor, pos: Tensor) -> Tensor:
    """(rows, H, L, L) bias for token positions `pos` (rows, L): base.py:586-617 (|i-j| * -slope),
    :305-308 (x clamp_min(alibi_scale, 0)), :681-698 (gather by ids_keep == evaluate at positions)."""
    slopes = torch.tensor(alibi_slopes(cfg.num_heads), dtype=torch.float32)
    dist = (pos[:, :, None] - pos[:, None, :]).abs().float()
    coef = slopes.view(1, -1, 1, 1) * scale.clamp_min(0).view(1, -1, 1, 1)
    return -coef * dist[:, None]


class LazyAlibi:
    """ALiBi bias evaluated per head / query chunk instead of as a (rows, H, L, L) tensor: the 48 kHz configuration
    (T = 12 000) would need 9.2 GB per copy. Same values as :func:`alibi_bias`."""

    def __init__(self, cfg: OracleConfig, scale: Tensor, pos: Tensor, chunk: int = 2048):
        self.coef = (torch.tensor(alibi_slopes(cfg.num_heads), dtype=torch.float32) * scale.clamp_min(0).view(-1))
        self.pos = pos.float()
        self.chunk = chunk

    def block(self, h: int, lo: int, hi: int) -> Tensor:
        """(rows, hi - lo, L) bias of head h for query rows lo..hi."""
        dist = (self.pos[:, lo:hi, None] - self.pos[:, None, :]).abs()
        return -self.coef[h] * dist


def attention(p: Dict[str, Tensor], pre: str, x: Tensor, bias, heads: int) -> Tensor:
    """AltAttention.forward (modules.py:368-410), dropout off. ``bias``: dense (rows, H, L, L) tensor or LazyAlibi."""
    b, n, c = x.shape
    qkv = F.linear(x, p[pre + "qkv.weight"], p[pre + "qkv.bias"]).reshape(b, n, 3, heads, c // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    if isinstance(bias, LazyAlibi):
        q = q * (c // heads) ** -0.5
        out = []
        for h in range(heads):
            rows = []
            for lo in range(0, n, bias.chunk):
                hi = min(n, lo + bias.chunk)
                a = (q[:, h, lo:hi] @ k[:, h].transpose(-2, -1)).float() + bias.block(h, lo, hi)
                rows.append(a.softmax(dim=-1) @ v[:, h])
            out.append(torch.cat(rows, 1))
        y = torch.stack(out, 2).reshape(b, n, c)
        return F.linear(y, p[pre + "proj.weight"], p[pre + "proj.bias"])
    attn = (q * (c // heads) ** -0.5) @ k.transpose(-2, -1)
    attn = (attn.float() + bias).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(y, p[pre + "proj.weight"], p[pre + "proj.bias"])


def alt_block(p: Dict[str, Tensor], pre: str, x: Tensor, bias: Tensor, cfg: OracleConfig) -> Tuple[Tensor, Tensor]:
    """AltBlock.forward post-LN branch (modules.py:328-337): returns (x, ffn target)."""
    d = cfg.embed_dim
    x = x + attention(p, pre + "attn.", x, bias, cfg.num_heads)
    r = x = F.layer_norm(x, (d,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], cfg.norm_eps)
    h = F.gelu(F.linear(x, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"]))
    t = F.linear(h, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
    x = F.layer_norm(r + t, (d,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], cfg.norm_eps)
    return x, t


def positional_encoder(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor) -> Tensor:
    """audio.py:93-113: depth x [grouped Conv1d(k, pad k//2), SamePad, LN(no affine), GELU]."""
    d, k = cfg.embed_dim, cfg.pos_kernel
    y = x.transpose(1, 2)
    for i in range(1, cfg.conv_pos_depth + 1):
        y = F.conv1d(y, p[ENC + f"relative_positional_encoder.{i}.0.weight"],
                     p[ENC + f"relative_positional_encoder.{i}.0.bias"], padding=k // 2, groups=cfg.conv_pos_groups)
        if k % 2 == 0:
            y = y[:, :, :-1]
        y = F.gelu(F.layer_norm(y.transpose(1, 2), (d,), None, None, 1e-5).transpose(1, 2))
    return y.transpose(1, 2)


def prenet(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, bias: Tensor) -> Tensor:
    """BlockEncoder.forward (modules.py:83-108), dropout off."""
    d = cfg.embed_dim
    x = F.layer_norm(x, (d,), p[ENC + "context_encoder.norm.weight"], p[ENC + "context_encoder.norm.bias"], cfg.norm_eps)
    for j in range(cfg.prenet_depth):
        x, _ = alt_block(p, ENC + f"context_encoder.blocks.{j}.", x, bias, cfg)
    return x


def decoder(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor) -> Tensor:
    """Decoder1d.forward (modules.py:179-192) incl. the residual rule (:124-134)."""
    dd, k = cfg.decoder_dim, cfg.decoder_kernel
    y = x.transpose(1, 2)
    residual = y
    for l in range(cfg.decoder_layers):
        y = F.conv1d(y, p[ENC + f"decoder.blocks.{l}.0.weight"], p[ENC + f"decoder.blocks.{l}.0.bias"], padding=k // 2,
                     groups=cfg.decoder_groups)
        y = F.gelu(F.layer_norm(y.transpose(1, 2), (dd,), None, None, 1e-5).transpose(1, 2))
        if residual.size(1) == y.size(1):
            y = y + residual
        residual = y
    return F.linear(y.transpose(1, 2), p[ENC + "decoder.proj.weight"], p[ENC + "decoder.proj.bias"])


def make_targets(layer_results: List[Tensor], k: int) -> Tensor:
    """data2vec2.py:1023-1066 with instance_norm_target_layer=True."""
    tl = [F.instance_norm(t.float().transpose(1, 2)).transpose(1, 2) for t in layer_results[-k:]]
    return sum(tl) / len(tl)


def compute_var(y: Tensor) -> Tensor:
    """data2vec2.py:1095-1110 (single process)."""
    y = y.view(-1, y.size(-1))
    return torch.sqrt(y.var(dim=0) + 1e-6).mean()


# ------------------------------------------------------------------------------------------------
# BC mixup (data2vec2.py:453-498, 536-598)
# ------------------------------------------------------------------------------------------------
def a_weight_table(fs: int, n_fft: int, min_db: float = -80.0) -> np.ndarray:
    freq = np.linspace(0, fs // 2, n_fft // 2 + 1)
    fsq = freq ** 2
    fsq[0] = 1.0
    w = 2.0 + 20.0 * (2 * np.log10(12194) + 2 * np.log10(fsq) - np.log10(fsq + 12194 ** 2) - np.log10(fsq + 20.6 ** 2)
                      - 0.5 * np.log10(fsq + 107.7 ** 2) - 0.5 * np.log10(fsq + 737.9 ** 2))
    return np.power(10, np.maximum(w, min_db) / 10)


def compute_gain(sound: Tensor, fs: int, wl: float, min_db: float = -80.0) -> Tensor:
    n_fft = round(fs * wl)
    aw = torch.from_numpy(a_weight_table(fs, n_fft, min_db))
    fr = sound.unfold(-1, n_fft, n_fft // 2)
    spec = torch.fft.rfft(torch.hann_window(n_fft) * fr)
    g = ((spec.abs() ** 2) * aw).sum(-1)
    return 10 * torch.log10(torch.maximum(g, torch.tensor(10 ** (min_db / 10))))


def mixup(cfg: OracleConfig, source: Tensor) -> Tensor:
    """Consumes the global torch CPU RNG exactly like the reference (same_mixup, mixup_prob=1):
    one uniform_ draw then one randperm."""
    r = torch.FloatTensor(1).uniform_(max(1e-6, cfg.source_mixup), 1).to(dtype=source.dtype)
    perm = torch.randperm(source.size(0))
    s2 = source[perm]
    g1, _ = compute_gain(source, cfg.sample_rate, cfg.mixing_window_length).max(-1)
    g1 = g1.to(source.dtype)
    g2 = g1[perm]
    pm = (1 / (1 + 10 ** ((g1 - g2) / 20) * (1 - r) / r)).unsqueeze(-1)
    return (pm * source + (1 - pm) * s2) / torch.sqrt(pm ** 2 + (1 - pm) ** 2)


# ------------------------------------------------------------------------------------------------
# the pretraining forward and the EMA step
# ------------------------------------------------------------------------------------------------
def annealed_decay(cfg: OracleConfig, num_updates: int) -> float:
    """data2vec2.py:396-405 + base.py:492-497."""
    if cfg.ema_decay == cfg.ema_end_decay:
        return cfg.ema_decay
    if num_updates >= cfg.ema_anneal_end_step:
        return cfg.ema_end_decay
    r = cfg.ema_end_decay - cfg.ema_decay
    return cfg.ema_end_decay - r * (1 - num_updates / cfg.ema_anneal_end_step)


def ema_step(student: Dict[str, Tensor], shadow: Dict[str, Tensor], decay: float) -> None:
    """fairseq EMAModule.step with ema_fp32=True, add_missing_params=False."""
    with torch.no_grad():
        for k, v in student.items():
            if k in shadow:
                shadow[k].mul_(decay).add_(v.detach().float(), alpha=1 - decay)


def pretrain_forward(student: Dict[str, Tensor], teacher: Dict[str, Tensor], cfg: OracleConfig, source: Tensor,
                     ids: Tensor, num_updates: int, *, do_mixup: bool = False, taps: Optional[dict] = None,
                     mask: Optional[np.ndarray] = None) -> Dict[str, Tensor]:
    """Data2VecMultiModel.forward pretraining branch (data2vec2.py:516-991) with every dropout
    and the decoder mask-token noise switched off (the deterministic "stage parity" setting of
    SURVEY.md section 7). Returns the reference's result dict (loss unreduced)."""
    d, m = cfg.embed_dim, cfg.clone_batch
    if do_mixup:
        with torch.no_grad():
            source = mixup(cfg, source)
        if taps is not None:
            taps["mixed_source"] = source
    lf = local_features(student, cfg, source, taps)  # (B, T, D)
    b, t, _ = lf.shape
    if mask is None:
        mask = pretrain_mask(cfg, b, t, ids, num_updates)
    mask_t = torch.from_numpy(mask)
    keep = ~mask_t
    tk = int(keep[0].sum())
    pos = torch.stack([torch.nonzero(keep[r]).flatten() for r in range(b * m)])  # ascending ids_keep
    x = lf.repeat_interleave(m, 0)
    x_masked = x * keep.unsqueeze(-1).to(x.dtype)  # encoder_zero_mask (base.py:464)
    x_pos = positional_encoder(student, cfg, x_masked)
    gidx = pos.unsqueeze(-1).expand(-1, -1, d)
    xs = torch.gather(x, 1, gidx) + torch.gather(x_pos, 1, gidx)  # base.py:278-280
    bias = alibi_bias(cfg, student[ENC + "alibi_scale"], pos)
    xs = prenet(student, cfg, xs, bias)
    if taps is not None:
        taps["student_prenet"] = torch.zeros(b * m, t, d).scatter_(1, gidx, xs.detach())
    for j in range(cfg.depth):
        xs, _ = alt_block(student, f"blocks.{j}.", xs, bias, cfg)
    if taps is not None:
        taps["student_out"] = torch.zeros(b * m, t, d).scatter_(1, gidx, xs.detach())
    dec_in = torch.zeros(b * m, t, d, dtype=xs.dtype).scatter(1, gidx, xs)  # mask tokens: std 0 -> zeros
    pred = decoder(student, cfg, dec_in)
    if taps is not None:
        taps["decoder_out"] = pred.detach()

    with torch.no_grad():
        tpos = torch.arange(t).unsqueeze(0).expand(b, -1)
        tbias = (LazyAlibi(cfg, teacher[ENC + "alibi_scale"], tpos) if t > 4096
                 else alibi_bias(cfg, teacher[ENC + "alibi_scale"], tpos))
        y = lf.detach() + positional_encoder(teacher, cfg, lf.detach())
        y = prenet(teacher, cfg, y, tbias)
        targets = []
        for j in range(cfg.depth):
            y, ffn = alt_block(teacher, f"blocks.{j}.", y, tbias, cfg)
            targets.append(ffn)
        y = make_targets(targets, cfg.average_top_k_layers)
    if taps is not None:
        taps["targets"] = y
        taps["local_features"] = lf.detach()
    yb = y.repeat_interleave(m, 0)[mask_t]
    xb = pred[mask_t]
    scale = cfg.loss_scale if cfg.loss_scale is not None else 1 / math.sqrt(d)
    loss = F.mse_loss(xb.float(), yb, reduction="none") * scale
    with torch.no_grad():
        res = {
            "losses": {"AUDIO_regression": loss},
            "sample_size": mask_t.sum().long(),
            "masked_pct": 1 - tk / t,
            "pred_var": compute_var(xb.float()),
            "target_var": compute_var(yb.float()),
            "mask": mask_t,
        }
    return res


def extract_features(student: Dict[str, Tensor], cfg: OracleConfig, source: Tensor) -> Dict[str, object]:
    """Data2VecMultiModel.extract_features (data2vec2.py:1112-1123 -> forward(features_only=True, mask=False),
    :632-728) in eval mode: the student on the unmasked full-length sequence. Returns the reference's dict
    (``layer_results`` = FFN outputs of the main blocks, what the finetune head averages, wav2vec2.py:446-462).
    First piece of the "next" row SURVEY.md section 8(f)-1; pinned by tests/golden/tiny_features.npz."""
    lf = local_features(student, cfg, source)
    b, t, _ = lf.shape
    pos = torch.arange(t).unsqueeze(0).expand(b, -1)
    bias = alibi_bias(cfg, student[ENC + "alibi_scale"], pos)
    x = lf + positional_encoder(student, cfg, lf)
    x = prenet(student, cfg, x, bias)
    layer_results = []
    for j in range(cfg.depth):
        x, ffn = alt_block(student, f"blocks.{j}.", x, bias, cfg)
        layer_results.append(ffn)
    return {"x": x, "linear_eval_projection": None, "padding_mask": None, "layer_results": layer_results, "mask": None}


def finetune_logits(student: Dict[str, Tensor], cfg: OracleConfig, source: Tensor, proj_w: Tensor, proj_b: Tensor,
                    top_k: Optional[int] = None) -> Tensor:
    """Wav2VecEncoderModOut.forward in eval mode (wav2vec2.py:437-482): extract_features, mean of the top-k FFN
    outputs, final dropout (off), ``proj`` = Linear(D, classes). Returns (B, T, classes) logits."""
    res = extract_features(student, cfg, source)
    k = top_k or cfg.average_top_k_layers
    lrs = res["layer_results"][-k:]
    x = sum(lrs) / len(lrs)
    return F.linear(x, proj_w, proj_b)


def finetune_features(student: Dict[str, Tensor], cfg: OracleConfig, source: Tensor, *,
                      time_mask: Optional[Tensor] = None, channel_mask: Optional[Tensor] = None) -> List[Tensor]:
    """Data2VecMultiModel.forward(features_only=True, mask=True) in TRAIN mode with every dropout, layerdrop and the
    mask-token noise at 0 (data2vec2.py:632-728; base.py:215-344 with clone_batch 1 / remove_masked False): the
    conv extractor runs without gradient (local_grad_mult 0, base.py:205-207; project_features still trains), masked frames are replaced by
    N(0, 0) = 0 tokens (encoder_zero_mask False, base.py:465-469), masked channels are zeroed (:470-484), then
    x + positional_encoder(x), prenet, main blocks. Returns the FFN outputs of the main blocks (with autograd graph).
    Pinned by tests/golden/tiny_finetune.npz (the reference's own forward + backward)."""
    lf = local_features(student, cfg, source, frozen_extractor=True)
    x = lf.clone()
    if time_mask is not None:
        x = x.masked_fill(time_mask[:, :, None], 0.0)
    if channel_mask is not None:
        x = x.masked_fill(channel_mask[:, None, :], 0.0)
    b, t, _ = x.shape
    pos = torch.arange(t).unsqueeze(0).expand(b, -1)
    bias = alibi_bias(cfg, student[ENC + "alibi_scale"], pos)
    x = x + positional_encoder(student, cfg, x)
    x = prenet(student, cfg, x, bias)
    layer_results = []
    for j in range(cfg.depth):
        x, ffn = alt_block(student, f"blocks.{j}.", x, bias, cfg)
        layer_results.append(ffn)
    return layer_results


def finetune_loss(student: Dict[str, Tensor], cfg: OracleConfig, source: Tensor, target: Tensor, proj_w: Tensor,
                  proj_b: Tensor, *, time_mask: Optional[Tensor] = None, channel_mask: Optional[Tensor] = None,
                  top_k: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """Wav2VecEncoderModOut.forward (wav2vec2.py:437-482, unfrozen phase, mixup off) + the focal branch of
    FinetuneCrossEntropyCriterion.forward (criterions.py:231-246, reduce=True): (summed loss, logits)."""
    lrs = finetune_features(student, cfg, source, time_mask=time_mask, channel_mask=channel_mask)
    k = top_k or cfg.average_top_k_layers
    top = lrs[-k:]
    logits = F.linear(sum(top) / len(top), proj_w, proj_b)
    return sigmoid_focal_loss(logits, target, reduction="sum"), logits


def sigmoid_focal_loss(inputs: Tensor, targets: Tensor, alpha: float = 0.25, gamma: float = 2.0,
                       reduction: str = "none") -> Tensor:
    """nn/utils.py:971-1010 (RetinaNet focal loss on logits, fp32): BCE-with-logits * (1 - p_t)^gamma * alpha_t."""
    x, t = inputs.float(), targets.float()
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def confusion_counts(logits: Tensor, targets: Tensor, threshold: float) -> Tuple[int, int, int, int]:
    """FinetuneCrossEntropyCriterion.compute_prec_rec_f1 (criterions.py:218-229) + confusion (utils.py:925-969),
    multi-label branch: predictions = sigmoid(logits) >= threshold, micro-summed over classes -> (tp, fp, tn, fn)."""
    pred = torch.sigmoid(logits.reshape(-1, logits.shape[-1]).float()) >= threshold
    tru = targets.reshape(-1, targets.shape[-1]) > 0.5
    tp = int((pred & tru).sum())
    fp = int((pred & ~tru).sum())
    tn = int((~pred & ~tru).sum())
    fn = int((~pred & tru).sum())
    return tp, fp, tn, fn


def make_teacher(student: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """make_ema_teacher / make_target_model (data2vec2.py:345-384): fp32 copy of the shared keys."""
    return {k: v.detach().clone().float() for k, v in student.items() if is_teacher_key(k)}


def pretrain_step(student: Dict[str, Tensor], teacher: Dict[str, Tensor], cfg: OracleConfig, source: Tensor,
                  ids: Tensor, num_updates: int, do_mixup: bool = False) -> Dict[str, object]:
    """forward + loss.sum().backward() + set_num_updates(num_updates + 1) (EMA step). The unit
    BASELINE.md section 4 times on the CPU."""
    for v in student.values():
        v.requires_grad_(True)
        v.grad = None
    res = pretrain_forward(student, teacher, cfg, source, ids, num_updates, do_mixup=do_mixup)
    loss = res["losses"]["AUDIO_regression"].sum()
    loss.backward()
    decay = annealed_decay(cfg, num_updates + 1)
    if decay < 1:
        ema_step(student, teacher, decay)
    return {"loss": loss.detach(), "sample_size": res["sample_size"], "ema_decay": decay, "result": res}
