"""Constructor surface of the reference's building blocks, same signatures and parameter names /
shapes / initialisation, for code that builds or inspects the model piecewise:

  SincConv                      nn/sinc.py:49-106            (forward runs the sm_100a sinc kernels)
  ConvFeatureExtractionModel    nn/utils.py:1044-1153
  AltBlock                      nn/modalities/modules.py:273-326
  Decoder1d                     nn/modalities/modules.py:137-178
  AudioEncoder                  nn/modalities/audio.py:57-149

The arithmetic of the pretraining / finetune step is NOT assembled from these modules' forwards -- it is the fused kernel
schedule of animal2vec_b200.engine / .finetune, reached through Data2VecMultiModel and Wav2VecCcasFinetune. Except for
SincConv, these classes hold parameters (their ``state_dict()`` matches the reference module's key for key) and raise
from ``forward`` with a pointer to the fused path; ``Data2VecMultiModel.load_state_dict`` accepts their tensors.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .config import D2vAudioConfig, D2vDecoderConfig, parse_conv_layers
from .params import sinc_buffers, sinc_mel_init

_FUSED = ("this module names and initialises parameters; the computation is the fused sm_100a kernel schedule reached "
          "through animal2vec_b200.data2vec2.Data2VecMultiModel / animal2vec_b200.wav2vec2.Wav2VecCcasFinetune")


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(_FUSED)


class SincConv(nn.Module):
    """nn/sinc.py:49-223 (non-learnable-kernel branch: band-pass filters from ``low_hz_`` / ``band_hz_``)."""

    def __init__(self, out_channels, kernel_size, input_shape=None, in_channels=1, stride=1, dilation=1, padding="same",
                 padding_mode="reflect", sample_rate=8000, min_low_hz=50, min_band_hz=None, learnable_filters=False,
                 apply_window_to_root=False, return_abs=False, init_scale="mel"):
        super().__init__()
        if learnable_filters or apply_window_to_root or return_abs:
            raise NotImplementedError("learnable_filters / apply_window_to_root / return_abs (not in the shipped recipe)")
        if in_channels != 1 or stride != 1 or dilation != 1 or padding != "same" or padding_mode != "reflect":
            raise NotImplementedError("SincConv variant other than 1 input channel, stride 1, reflect 'same' padding")
        if init_scale != "mel":
            raise NotImplementedError("init_scale other than 'mel'")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.dilation, self.padding, self.padding_mode = stride, dilation, padding, padding_mode
        self.sample_rate, self.min_low_hz = sample_rate, min_low_hz
        self.min_band_hz = int(np.ceil(sample_rate / kernel_size)) if min_band_hz is None else min_band_hz
        low, band = sinc_mel_init(out_channels, kernel_size, sample_rate)
        self.low_hz_ = nn.Parameter(low)
        self.band_hz_ = nn.Parameter(band)
        n_, window = sinc_buffers(kernel_size, sample_rate)
        self.register_buffer("n_", n_, persistent=False)
        self.register_buffer("window_", window, persistent=False)

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """(B, 1, N) or (B, N) -> (B, out_channels, N) fp32 (forward only, through a2v_sinc_filters_fwd / a2v_sinc_conv_fwd)."""
        from . import ops

        x = waveforms.squeeze(1) if waveforms.dim() == 3 else waveforms
        x = x.to(torch.float32).contiguous()
        filt = ops.sinc_filters_fwd(self.low_hz_.detach().view(-1), self.band_hz_.detach().view(-1), self.n_, self.window_,
                                    self.kernel_size, float(self.min_low_hz), float(self.min_band_hz), float(self.sample_rate))
        y = ops.sinc_conv_fwd(x, filt, torch.float32)  # (B, N, 128) channels-last, column 127 is padding
        return y[..., : self.out_channels].transpose(1, 2)


class PSwish(_ParamsOnly):
    """nn/utils.py:1413-1435."""

    def __init__(self, num_features: int):
        super().__init__()
        self.p_swish_alpha = nn.Parameter(torch.full((1, num_features, 1), 2.0))
        self.p_swish_beta = nn.Parameter(torch.zeros(1, num_features, 1))


def _ln(dim: int, affine: bool = True) -> nn.LayerNorm:
    return nn.LayerNorm(dim, elementwise_affine=affine)


class ConvFeatureExtractionModel(_ParamsOnly):
    """nn/utils.py:1044-1153 for mode='layer_norm' (+ sinc_input / sinc_norm='layer_norm' / use_pswish)."""

    def __init__(self, conv_layers, dropout: float = 0.0, mode: str = "default", conv_bias: bool = False,
                 sinc_input: bool = False, apply_window_to_root: bool = False, sample_rate=8000, sinc_norm="layer_norm",
                 use_pswish=False):
        super().__init__()
        if mode != "layer_norm" or conv_bias or apply_window_to_root or dropout:
            raise NotImplementedError("extractor variant other than mode='layer_norm', no bias, no dropout")
        if sinc_input and sinc_norm != "layer_norm":
            raise NotImplementedError("sinc_norm other than 'layer_norm' (pcen / instance are not on the shipped path)")
        self.conv_layers = nn.ModuleList()
        in_d = 1
        for i, (dim, k, stride) in enumerate(conv_layers):
            if sinc_input and i == 0:
                conv = SincConv(out_channels=dim, kernel_size=k, stride=stride, sample_rate=sample_rate)
                act = PSwish(dim) if use_pswish else nn.GELU()
            else:
                conv = nn.Conv1d(in_d, dim, k, stride=stride, bias=False,
                                 padding="same" if stride == 1 else int(np.ceil(stride / 2)))
                nn.init.kaiming_normal_(conv.weight)
                act = nn.GELU()
            # Sequential(conv, Dropout, Sequential(TransposeLast, Fp32LayerNorm, TransposeLast), activation)
            self.conv_layers.append(nn.Sequential(conv, nn.Dropout(p=dropout),
                                                  nn.Sequential(nn.Identity(), _ln(dim), nn.Identity()), act))
            in_d = dim


class AltAttention(_ParamsOnly):
    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_ParamsOnly):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class AltBlock(_ParamsOnly):
    """nn/modalities/modules.py:273-326."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 mlp_drop=0.0, post_mlp_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 layer_norm_first=True, ffn_targets=False, cosine_attention=False):
        super().__init__()
        if cosine_attention or qk_scale is not None:
            raise NotImplementedError("cosine_attention / qk_scale")
        self.layer_norm_first, self.ffn_targets = layer_norm_first, ffn_targets
        self.norm1 = norm_layer(dim)
        self.attn = AltAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class Decoder1d(_ParamsOnly):
    """nn/modalities/modules.py:137-178."""

    def __init__(self, cfg: D2vDecoderConfig, input_dim):
        super().__init__()
        if cfg.projection_layers != 1:
            raise NotImplementedError("decoder projection_layers != 1")
        self.decoder_cfg = cfg
        blocks = []
        for i in range(cfg.decoder_layers):
            conv = nn.Conv1d(input_dim if i == 0 else cfg.decoder_dim, cfg.decoder_dim, kernel_size=cfg.decoder_kernel,
                             padding=cfg.decoder_kernel // 2, groups=cfg.decoder_groups)
            # Sequential(conv, SamePad, TransposeLast, LayerNorm(no affine), TransposeLast, GELU)
            blocks.append(nn.Sequential(conv, nn.Identity(), nn.Identity(), _ln(cfg.decoder_dim, False), nn.Identity(), nn.GELU()))
        self.blocks = nn.Sequential(*blocks)
        self.proj = nn.Linear(cfg.decoder_dim, input_dim)


class BlockEncoder(_ParamsOnly):
    def __init__(self, blocks, norm_layer, layer_norm_first, layerdrop, dropout):
        super().__init__()
        self.blocks = blocks
        self.norm = norm_layer
        self.layer_norm_first, self.layerdrop = layer_norm_first, layerdrop
        self.dropout = nn.Dropout(dropout, inplace=True)


class AudioEncoder(_ParamsOnly):
    """nn/modalities/audio.py:57-149 (+ the alibi_scale parameter of nn/modalities/base.py:116-134)."""

    def __init__(self, modality_cfg: D2vAudioConfig, embed_dim: int, make_block: Callable[[float], nn.Module],
                 norm_layer: Callable[[int], nn.LayerNorm], layer_norm_first: bool, alibi_biases: Optional[Dict] = None,
                 task=None):
        super().__init__()
        a = modality_cfg
        self.modality_cfg = a
        self.feature_enc_layers = parse_conv_layers(a.conv_feature_layers)
        c_last = self.feature_enc_layers[-1][0]
        self.local_encoder = ConvFeatureExtractionModel(
            conv_layers=self.feature_enc_layers, dropout=0.0, mode=a.extractor_mode, conv_bias=False,
            sinc_input=a.sinc_input, apply_window_to_root=a.apply_window_to_root, sample_rate=a.sample_rate,
            sinc_norm=a.sinc_norm, use_pswish=a.use_pswish)
        self.project_features = nn.Sequential(nn.Identity(), _ln(c_last), nn.Linear(c_last, embed_dim))
        k = max(3, a.conv_pos_width // a.conv_pos_depth)
        pos = [nn.Identity()]  # TransposeLast
        for _ in range(a.conv_pos_depth):
            conv = nn.Conv1d(embed_dim, embed_dim, kernel_size=k, padding=k // 2, groups=a.conv_pos_groups)
            # Sequential(conv, SamePad, TransposeLast, LayerNorm(no affine), TransposeLast, GELU)
            pos.append(nn.Sequential(conv, nn.Identity(), nn.Identity(), _ln(embed_dim, False), nn.Identity(), nn.GELU()))
        pos.append(nn.Identity())
        self.relative_positional_encoder = nn.Sequential(*pos)
        dpr = np.linspace(a.start_drop_path_rate, a.end_drop_path_rate, a.prenet_depth)
        self.context_encoder = BlockEncoder(nn.ModuleList(make_block(float(dpr[i])) for i in range(a.prenet_depth)),
                                            norm_layer(embed_dim) if not layer_norm_first else None, layer_norm_first,
                                            a.prenet_layerdrop, a.prenet_dropout)
        self.decoder = Decoder1d(a.decoder, embed_dim) if a.decoder is not None else None
        if a.use_alibi_encoder:
            heads = a.num_alibi_heads if a.learned_alibi_scale_per_head else 1
            layers = (a.prenet_depth + a.model_depth) if a.learned_alibi_scale_per_layer else 1
            self.alibi_scale = nn.Parameter(torch.full((layers, 1, heads, 1, 1), float(a.alibi_scale)),
                                            requires_grad=a.learned_alibi_scale)
        else:
            self.alibi_scale = None


def reference_state_keys(module: nn.Module) -> Dict[str, tuple]:
    """state_dict keys -> shapes of a constructor-surface module (parameters only)."""
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}
