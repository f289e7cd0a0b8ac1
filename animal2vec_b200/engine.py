"""The pretraining step (forward, backward, EMA) as an explicit schedule of sm_100a kernel launches.

This is the host side of the hot path: it owns device buffers (through torch) and decides which
C-ABI kernel runs when; it performs no arithmetic of its own. Stage by stage it replaces
  nn/data2vec2.py:516-991 (Data2VecMultiModel.forward, pretraining branch) and its autograd backward,
  nn/modalities/base.py:194-344 (local_features / contextualized_features), :162-192 (decoder_input),
  nn/modalities/audio.py:57-149 (feature extractor, projection, positional conv, prenet, decoder),
  nn/modalities/modules.py:74-108,137-192,272-410 (BlockEncoder, Decoder1d, AltBlock, AltAttention),
  nn/sinc.py:107-223 and nn/utils.py:1043-1163 (SincConv + ConvFeatureExtractionModel),
  nn/data2vec2.py:386-410 (set_num_updates -> EMA teacher step)
of /root/reference. Activations are channels-last everywhere; bf16 with fp32 statistics in the
production mode, fp32 (GEMMs through a 3-term bf16 hi/lo split on the same tcgen05 kernel) in the
validation mode used for the 1e-3 per-stage parity gate.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import gemm, masking, ops
from . import lib as L
from . import params as P
from .config import Data2VecMultiConfig, parse_conv_layers, resolve

ENC = P.ENC
GP = 64  # stored width of one channel group in the group-padded decoder layout


def alibi_slopes(n: int) -> List[float]:
    """nn/modalities/base.py:559-576 (geometric head slopes, interleaved for non powers of two)."""
    def pow2(m):
        start = 2 ** (-(2 ** -(math.log2(m) - 3)))
        return [start * start ** i for i in range(m)]

    if math.log2(n).is_integer():
        return pow2(n)
    c = 2 ** math.floor(math.log2(n))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n - c]


def a_weight_table(fs: int, n_fft: int, min_db: float = -80.0) -> np.ndarray:
    """nn/data2vec2.py:461-479: 10^(A-weighting dB / 10) per rfft bin, float64."""
    freq = np.linspace(0, fs // 2, n_fft // 2 + 1)
    fsq = freq ** 2
    fsq[0] = 1.0
    w = 2.0 + 20.0 * (2 * np.log10(12194) + 2 * np.log10(fsq) - np.log10(fsq + 12194 ** 2)
                      - np.log10(fsq + 20.6 ** 2) - 0.5 * np.log10(fsq + 107.7 ** 2)
                      - 0.5 * np.log10(fsq + 737.9 ** 2))
    return np.power(10, np.maximum(w, min_db) / 10)


def annealed_decay(cfg: Data2VecMultiConfig, num_updates: int) -> float:
    """nn/data2vec2.py:396-405 + nn/modalities/base.py:492-497 (linear anneal)."""
    if cfg.ema_decay == cfg.ema_end_decay:
        return cfg.ema_decay
    if num_updates >= cfg.ema_anneal_end_step:
        return cfg.ema_end_decay
    r = cfg.ema_end_decay - cfg.ema_decay
    return cfg.ema_end_decay - r * (1 - num_updates / cfg.ema_anneal_end_step)


class _Weights:
    """Operand lookup for one model (student or teacher)."""

    def __init__(self):
        self.fwd: Dict[str, torch.Tensor] = {}    # name -> bf16 NT operand
        self.dgrad: Dict[str, torch.Tensor] = {}  # name -> bf16 operand of the data-gradient product
        self.split: Dict[str, int] = {}           # name -> K granularity of the fp32-mode split (fwd)
        self.split_d: Dict[str, int] = {}
        self.dgrad_s: Dict[str, torch.Tensor] = {}  # name -> operand of the im2col-free strided data gradient
        self.f32: Dict[str, torch.Tensor] = {}    # name -> fp32 view / padded fp32 buffer (biases, norms)
        self.alibi: Optional[torch.Tensor] = None


class PretrainEngine:
    def __init__(self, cfg: Data2VecMultiConfig, device="cuda", precision: str = "bf16",
                 init: Optional[Dict[str, torch.Tensor]] = None, init_seed: int = 0, rng_seed: int = 0,
                 finetune: bool = False):
        cfg = resolve(cfg)
        self._check_supported(cfg, finetune=finetune)
        self.finetune = finetune
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        L.load()  # fail loudly if the CUDA library is not built: there is no other implementation
        self.cfg = cfg
        self.device = torch.device(device)
        self.fp32 = precision == "fp32"
        self.adt = torch.float32 if self.fp32 else torch.bfloat16
        self.rng_seed = rng_seed
        self.step_counter = 0
        a = cfg.modalities.audio
        self.a = a
        self.D = cfg.embed_dim
        self.H = cfg.num_heads
        self.M = cfg.clone_batch
        self.hidden = int(cfg.embed_dim * cfg.mlp_ratio)
        self.layers = parse_conv_layers(a.conv_feature_layers)
        self.kp = P.pos_kernel(cfg)
        self.dec = a.decoder
        if self.D % 64 or self.D // self.H != 64:
            raise NotImplementedError("the attention kernels are built for head_dim 64")

        shapes, order, shared = P.student_layout(cfg)
        self.S = P.FlatParams(shapes, self.device, with_grad=True, order=order)
        self.shared_names = shared
        self.E = P.FlatParams({k: shapes[k] for k in shared}, self.device, with_grad=False, order=shared)
        assert all(self.E.offsets[k] == self.S.offsets[k] for k in shared)
        self.S16 = torch.zeros(self.S.total, device=self.device, dtype=torch.bfloat16)
        self.T16 = torch.zeros(self.E.total, device=self.device, dtype=torch.bfloat16)
        self.load_student(init if init is not None else P.default_init(cfg, init_seed))
        self.reset_teacher()

        # constant buffers
        c0, k0, _ = self.layers[0]
        n_, window = P.sinc_buffers(k0, a.sample_rate)
        self.sinc_n = n_.to(self.device)
        self.sinc_window = window.to(self.device)
        self.min_low_hz = 50.0
        self.min_band_hz = float(math.ceil(a.sample_rate / k0))  # nn/sinc.py:79
        self.slopes = torch.tensor(alibi_slopes(self.H), dtype=torch.float32, device=self.device)
        if cfg.source_mixup >= 0:
            n_fft = round(cfg.sample_rate * cfg.mixing_window_length)
            self.n_fft = n_fft
            self.hann = torch.hann_window(n_fft).float().to(self.device)
            self.aweight = torch.from_numpy(a_weight_table(cfg.sample_rate, n_fft)).float().to(self.device)

        self._pack_table_s = self._pack_table_t = self._unpack_table = None
        # full-length (teacher) attention visits only the key tiles ALiBi leaves above 2^-50 (see ops.attn_fwd)
        self.alibi_locality = True
        # the im2col operands of the strided feature-extractor convs (85 MB per clip) are kept for the weight gradients
        self.keep_im2col = True
        self._build_packs()
        self.ctx = None
        self.kernel_launches = 0
        self.has_teacher = True
        self.has_decoder = True
        # bf16: the last positional-conv layer of the student runs on the kept rows only (see _posconv_forward)
        self.sparse_last_posconv = True

    # ------------------------------------------------------------------------------------ support matrix
    @staticmethod
    def _check_supported(cfg: Data2VecMultiConfig, finetune: bool = False) -> None:
        """``finetune``: the variants the finetune wrapper switches on through its arg_overrides
        (nn/wav2vec2.py:95-130: layerdrop, activation dropout, noise mask tokens, channel masking, frozen feature
        extractor) are accepted; they are implemented by animal2vec_b200.finetune.FinetuneEngine only."""
        a = cfg.modalities.audio
        bad = []
        if cfg.layer_norm_first: bad.append("layer_norm_first=True")
        if (cfg.layerdrop or a.prenet_layerdrop) and not finetune: bad.append("layerdrop>0")
        if cfg.start_drop_path_rate or cfg.end_drop_path_rate or a.start_drop_path_rate or a.end_drop_path_rate:
            bad.append("drop_path>0")
        if cfg.activation_dropout and not finetune: bad.append("activation_dropout>0")
        if cfg.dropout_input: bad.append("dropout_input>0")
        if cfg.end_of_block_targets: bad.append("end_of_block_targets")
        if not cfg.instance_norm_target_layer or cfg.layer_norm_target_layer or cfg.batch_norm_target_layer:
            bad.append("target layer norm other than instance_norm_target_layer")
        if cfg.instance_norm_targets or cfg.layer_norm_targets: bad.append("instance/layer_norm_targets")
        if cfg.loss_beta != 0: bad.append("loss_beta>0 (smooth L1)")
        if cfg.cls_loss or cfg.recon_loss: bad.append("cls_loss/recon_loss")
        if cfg.ema_encoder_only: bad.append("ema_encoder_only=True")
        if cfg.shared_decoder is not None: bad.append("shared_decoder")
        if not a.sinc_input or a.sinc_norm != "layer_norm" or not a.use_pswish or a.extractor_mode != "layer_norm":
            bad.append("feature extractor other than sinc_input + layer_norm + PSwish")
        if a.apply_window_to_root: bad.append("apply_window_to_root")
        if not a.use_alibi_encoder or a.learned_alibi or a.learned_alibi_scale_per_layer:
            bad.append("ALiBi variant other than the (learned per-head / global) scale")
        if a.num_extra_tokens: bad.append("num_extra_tokens>0")
        if a.inverse_mask or a.keep_masked_pct or a.remove_masks or a.mask_prob_min is not None:
            bad.append("mask variant")
        if a.mask_channel_prob and not finetune: bad.append("mask_channel_prob>0")
        if a.mask_length == 1: bad.append("mask_length=1 (random_masking branch, base.py:394-395)")
        if not a.encoder_zero_mask and not finetune: bad.append("encoder_zero_mask=False")
        if a.ema_local_encoder: bad.append("ema_local_encoder")
        if a.local_grad_mult != (0.0 if finetune else 1.0):
            bad.append("local_grad_mult other than 1 (pretraining) / 0 (finetune, feature_grad_mult: 0.0)")
        if a.conv_pos_pre_ln: bad.append("conv_pos_pre_ln")
        d = a.decoder
        if d is None or d.add_positions_masked or d.add_positions_all or d.projection_layers != 1 or not d.decoder_residual:
            bad.append("decoder variant")
        if cfg.source_mixup >= 0 and (not cfg.same_mixup or cfg.mixup_prob < 1 or cfg.gain_mode != "A_weighting"):
            bad.append("mixup variant other than same_mixup / mixup_prob=1 / A_weighting")
        if bad:
            raise NotImplementedError("not on the shipped pretraining path (SURVEY.md section 8): " + ", ".join(bad))

    # ------------------------------------------------------------------------------------ parameters
    def load_student(self, tensors: Dict[str, torch.Tensor]) -> None:
        for k in self.S.names:
            if k not in tensors:
                raise KeyError(f"missing parameter {k}")
            self.S.view(k).copy_(tensors[k].to(self.device, torch.float32).reshape(self.S.shapes[k]))
        self._student_dirty = True
        self._s16_valid = False

    def reset_teacher(self) -> None:
        """make_ema_teacher (nn/data2vec2.py:345-384): fp32 copy of the shared student parameters."""
        self.E.data.copy_(self.S.data[: self.E.total])
        self._teacher_dirty = True
        self._t16_valid = False

    def load_teacher(self, tensors: Dict[str, torch.Tensor]) -> None:
        for k in self.E.names:
            self.E.view(k).copy_(tensors[k].to(self.device, torch.float32).reshape(self.E.shapes[k]))
        self._teacher_dirty = True
        self._t16_valid = False

    def drop_teacher(self) -> None:
        """remove_pretraining_modules (nn/data2vec2.py:1125-1127): the EMA teacher and its fp32 shadow are released."""
        self.has_teacher = False
        self.E.data = self.E.data[:0]
        self.T16 = self.T16[:0]
        self.WT = _Weights()
        self._pack_table_t = None

    def drop_decoder(self) -> None:
        """remove_pretraining_modules(keep_decoder=False): the decoder parameters stay in the flat buffer (unused,
        zero gradient) but leave the state dict; pretraining forwards are refused from here on."""
        self.has_decoder = False

    def mark_student_updated(self, s16_valid: bool = False) -> None:
        """The fp32 masters changed (optimizer step / state-dict load). ``s16_valid``: the bf16 flat copy was
        already refreshed by the writer (the fused AdamW kernel stores it)."""
        self._student_dirty = True
        self._s16_valid = s16_valid

    def _lin_names(self, prefix_list: Sequence[str]) -> List[str]:
        out = []
        for pre in prefix_list:
            out += [pre + "attn.qkv.weight", pre + "attn.proj.weight", pre + "mlp.fc1.weight", pre + "mlp.fc2.weight"]
        return out

    def _build_packs(self) -> None:
        cfg, a, d = self.cfg, self.a, self.D
        self.block_prefixes = [ENC + f"context_encoder.blocks.{j}." for j in range(a.prenet_depth)] + \
                              [f"blocks.{j}." for j in range(cfg.depth)]
        self.lin_names = self._lin_names(self.block_prefixes)
        sp: Dict[str, P.Pack] = {}
        # transposes of every student Linear for the data gradients
        for n in self.lin_names + [ENC + "project_features.2.weight"]:
            no, ki = self.S.shapes[n]
            sp[n + "|T"] = P.pack_linear_t(n, no, ki)
        # feature-extractor convs (layer 0 is the sinc front end)
        cin, cinp = self.layers[0][0], 128
        le = ENC + "local_encoder.conv_layers."
        for i, (c, k, st) in enumerate(self.layers[1:], start=1):
            n = le + f"{i}.0.weight"
            sp[n + "|F"] = P.pack_conv_fwd(n, 1, c, cin, k, cgp=cinp)
            if st > 1:
                sp[n + "|T"] = P.pack_col_t(n, c, cin, k, cinp)  # im2col fallback (fp32 mode, odd clip lengths)
                if not self.fp32:  # im2col-free strided path (gemm.strided_conv_*)
                    sp[n + "|S"] = P.pack_conv_dgrad_strided(n, c, cin, k, st, int(math.ceil(st / 2)), cinp)
            else:
                sp[n + "|T"] = P.pack_conv_dgrad(n, 1, c, cin, k, cgp=cinp)
            cin, cinp = c, c
        # positional convs
        g = a.conv_pos_groups
        self.pos_names = [ENC + f"relative_positional_encoder.{i}.0.weight" for i in range(1, a.conv_pos_depth + 1)]
        tp: Dict[str, P.Pack] = {}
        for n in self.pos_names:
            sp[n + "|F"] = P.pack_conv_fwd(n, g, d // g, d // g, self.kp)
            sp[n + "|T"] = P.pack_conv_dgrad(n, g, d // g, d // g, self.kp)
            tp[n + "|F"] = P.pack_conv_fwd(n, g, d // g, d // g, self.kp)
        # decoder (group-padded: dd/groups real channels in GP-wide groups)
        dc = self.dec
        dg = dc.decoder_groups
        self.dec_ng = dc.decoder_dim // dg
        if self.dec_ng > GP or (d // dg) % 64 or (d // a.conv_pos_groups) % 64 or dc.decoder_dim == d:
            raise NotImplementedError("channel-group widths: embed_dim/groups must be a multiple of 64, decoder "
                                      "groups at most 64 wide and decoder_dim != embed_dim")
        # bf16: compact storage (dec_ng channels per group, nothing between the groups) -- the slab kernels address the
        # groups dec_ng apart and skip the K steps over what would be padding; the decoder layers are HBM bound, the
        # padding was a quarter of their bytes. The weight packs keep GP-wide K / N blocks either way. fp32 validation
        # mode: GP-wide groups (its tap-loop GEMM needs 64-channel groups).
        self.dec_gw = self.dec_ng if (not self.fp32 and self.dec_ng == 48) else GP
        self.dec_wp = dg * self.dec_gw  # stored decoder width
        for l in range(dc.decoder_layers):
            n = ENC + f"decoder.blocks.{l}.0.weight"
            cg = (d // dg) if l == 0 else self.dec_ng
            cgp = (d // dg) if l == 0 else GP
            sp[n + "|F"] = P.pack_conv_fwd(n, dg, self.dec_ng, cg, dc.decoder_kernel, ngp=GP, cgp=cgp)
            sp[n + "|T"] = P.pack_conv_dgrad(n, dg, self.dec_ng, cg, dc.decoder_kernel, ngp=GP, cgp=cgp)
            sp[ENC + f"decoder.blocks.{l}.0.bias|B"] = P.pack_bias_padded(ENC + f"decoder.blocks.{l}.0.bias", dg,
                                                                          self.dec_ng, self.dec_gw)
        n = ENC + "decoder.proj.weight"
        sp[n + "|F"] = P.pack_cols_padded(n, d, dg, self.dec_ng, self.dec_gw)
        sp[n + "|T"] = P.pack_cols_padded_t(n, d, dg, self.dec_ng, self.dec_gw)
        self.sp, self.tp = sp, tp
        # packed fp32 gradient buffers for weights whose GEMM layout differs from the checkpoint layout
        # (tap convs: TRANSPOSED layout (G*k*Cg, Ng) written by gemm.conv_wgrad_tn, see params.pack_conv_fwd)
        self.gpacked: Dict[str, torch.Tensor] = {}
        self.gt_keys = set()
        for i, (_c, _k, st) in enumerate(self.layers[1:], start=1):
            if st == 1 or not self.fp32:  # bf16: the strided layers' weight gradients use the transposed layout too
                self.gt_keys.add(le + f"{i}.0.weight|F")
        for n in self.pos_names:
            self.gt_keys.add(n + "|F")
        for l in range(dc.decoder_layers):
            self.gt_keys.add(ENC + f"decoder.blocks.{l}.0.weight|F")
        for key, pk in sp.items():
            if key.endswith("|F") or key.endswith("|B"):
                shape = pk.gt_shape if key in self.gt_keys else pk.out_shape
                self.gpacked[key] = torch.zeros(shape, device=self.device, dtype=torch.float32)
        self.WS, self.WT = _Weights(), _Weights()

    def _refresh_student(self) -> None:
        if not self._student_dirty:
            return
        S, W = self.S, self.WS
        if not self.fp32 and not self._s16_valid:
            ops.cast_bf16(S.data, out=self.S16)
            self._s16_valid = True
        for n in self.lin_names + [ENC + "project_features.2.weight"]:
            if self.fp32:
                W.fwd[n] = ops.split3(S.view(n), 1)
            else:
                W.fwd[n] = S.view(n, self.S16)
            W.split[n] = S.shapes[n][1]
        if not self.fp32 and self._pack_table_s is not None:
            # every operand buffer is persistent: one table-driven launch rebuilds all the packs
            self._pack_table_s.run()
            self._student_dirty = False
            return
        table = None if self.fp32 else ops.RelayoutTable(self.device)
        for key, pk in self.sp.items():
            name, kind = key.split("|")
            buf = P.materialize(pk, S.view(name), self.fp32)
            if table is not None:
                for dims, ist, ioff, ost, ooff in pk.parts:
                    table.add(S.view(name), buf, dims, ist, ioff, ost, ooff)
            if kind == "F":
                W.fwd[name], W.split[name] = buf, pk.split_k
            elif kind == "T":
                W.dgrad[name], W.split_d[name] = buf, pk.split_k
            elif kind == "S":
                W.dgrad_s[name] = buf
            else:
                W.f32[name] = buf
        for n in S.names:
            if n not in W.f32:
                W.f32[n] = S.view(n)
        W.alibi = self._alibi_of(S, self.WS)
        self._pack_table_s = table
        self._student_dirty = False

    def _alibi_of(self, F: P.FlatParams, W: _Weights) -> torch.Tensor:
        # the parameter always exists (base.py:116-134); learned_alibi_scale only decides whether it gets a gradient
        return F.view(ENC + "alibi_scale").view(-1)

    def _refresh_teacher(self) -> None:
        if not self._teacher_dirty:
            return
        E, W = self.E, self.WT
        if not self.fp32 and not self._t16_valid:
            ops.cast_bf16(E.data, out=self.T16)
            self._t16_valid = True
        for n in self.lin_names:
            if self.fp32:
                W.fwd[n] = ops.split3(E.view(n), 1)
            else:
                W.fwd[n] = E.view(n, self.T16)
            W.split[n] = E.shapes[n][1]
        if not self.fp32 and self._pack_table_t is not None:
            self._pack_table_t.run()
            self._teacher_dirty = False
            return
        table = None if self.fp32 else ops.RelayoutTable(self.device)
        for key, pk in self.tp.items():
            name, _ = key.split("|")
            W.fwd[name], W.split[name] = P.materialize(pk, E.view(name), self.fp32), pk.split_k
            if table is not None:
                table.add(E.view(name), W.fwd[name], pk.dims, pk.in_strides, pk.in_off, pk.out_strides, 0)
        for n in E.names:
            W.f32[n] = E.view(n)
        W.alibi = self._alibi_of(E, W)
        self._pack_table_t = table
        self._teacher_dirty = False

    # ------------------------------------------------------------------------------------ GEMM helpers
    def _split_a(self, a: torch.Tensor, k: int) -> torch.Tensor:
        return ops.split3(a.reshape(-1, k), 0).view(*a.shape[:-1], 3 * a.shape[-1])

    def lin(self, a, W: _Weights, name, *, dgrad=False, **kw):
        w = (W.dgrad if dgrad else W.fwd)[name]
        if self.fp32:
            a = self._split_a(a, (W.split_d if dgrad else W.split)[name])
        return gemm.gemm_nt(a, w, out_dtype=self.adt, **kw)

    def conv(self, x, W: _Weights, name, *, taps, pad, groups, dgrad=False, bias=None, x_real=0, ng_out=None):
        """``x_real``: channels per input group that can be non-zero when that is less than the 64 K columns the
        weights hold per tap (decoder: dec_ng): the slab kernel skips the K steps over the rest. ``ng_out``: outputs
        per group to write, that many columns apart (compact decoder layout)."""
        w = (W.dgrad if dgrad else W.fwd)[name]
        if self.fp32:
            x = self._split_a(x, (W.split_d if dgrad else W.split)[name])
        elif gemm.conv_slab_ok(x, w, taps, groups):
            return gemm.conv_slab(x, w, taps=taps, pad=pad, groups=groups, out_dtype=self.adt, bias=bias,
                                  x_real_cols=x_real, ng_out=ng_out)
        assert ng_out is None or ng_out == w.shape[0] // groups
        return gemm.conv_nt(x, w, taps=taps, pad=pad, groups=groups, out_dtype=self.adt, bias=bias)

    def wgrad(self, dy, x, out):
        """out (M, N) fp32 += dy^T x."""
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        if self.fp32:
            dy2, x2 = ops.split3(dy2, 2), ops.split3(x2, 3)
        gemm.gemm_tn(dy2, x2, out.view(dy2.shape[-1], x2.shape[-1]))

    def conv_wgrad(self, dy, x, out, *, taps, pad, groups):
        if self.fp32:
            b, t, n = dy.shape
            dy = ops.split3(dy.reshape(-1, n), 2).view(3 * b, t, n)
            x = ops.split3(x.reshape(-1, x.shape[-1]), 3).view(3 * b, t, x.shape[-1])
        elif x.shape[-1] in (groups * 64, groups * 48) and dy.shape[-1] // groups <= 64 and taps <= 32:
            gemm.conv_slab_wgrad(dy, x, out, taps=taps, pad=pad, groups=groups)
            return
        gemm.conv_wgrad_tn(dy, x, out, taps=taps, pad=pad, groups=groups)

    def _seed(self, site: int) -> int:
        return ((self.rng_seed * 1000003 + self.step_counter) * 4096 + site) & 0xFFFFFFFFFFFFFFFF

    def G(self, name: str) -> torch.Tensor:
        return self.S.gview(name)

    # ------------------------------------------------------------------------------------ stages: forward
    def _mixup(self, x: torch.Tensor, info: Optional[dict] = None) -> torch.Tensor:
        """nn/data2vec2.py:536-598 (same_mixup, mixup_prob = 1, A-weighted gain). The random draws come
        from torch's CPU generator in the reference's order: one uniform_ for r, then one randperm."""
        cfg = self.cfg
        r = float(torch.FloatTensor(1).uniform_(max(1e-6, cfg.source_mixup), 1).item())
        perm = ops.h2d_async(torch.randperm(x.size(0)).to(torch.int32), self.device)
        gain = ops.mixup_gain(x, self.hann, self.aweight, self.n_fft, self.n_fft // 2)
        mixed, _ = ops.mixup_apply(x, perm, gain, r)
        if info is not None:  # the finetune path mixes the targets with the same draw (nn/wav2vec2.py:424-431)
            info["perm"], info["r"] = perm, r
        return mixed

    def _fe_forward(self, x: torch.Tensor, c: SimpleNamespace, save: bool, save_proj: bool = False) -> torch.Tensor:
        """SincConv + conv stack + project_features -> local_features (B, T, D). ``save_proj`` (with save False):
        keep only what the backward of project_features needs (frozen conv extractor, local_grad_mult = 0:
        base.py:205-207 puts just ``local_encoder`` under no_grad, the projection still trains)."""
        W = self.WS
        le = ENC + "local_encoder.conv_layers."
        c0, k0, _ = self.layers[0]
        filt = ops.sinc_filters_fwd(W.f32[le + "0.0.low_hz_"].view(-1), W.f32[le + "0.0.band_hz_"].view(-1),
                                    self.sinc_n, self.sinc_window, k0, self.min_low_hz, self.min_band_hz,
                                    float(self.a.sample_rate))
        y = ops.sinc_conv_fwd(x, filt, self.adt)  # (B, N, 128)
        cfg0 = ops.RowLnCfg(128, 1e-5, act=2, group_width=128, group_real=c0)
        act, m, r = ops.rowln_fwd(cfg0, y, None, W.f32[le + "0.2.1.weight"], W.f32[le + "0.2.1.bias"],
                                  W.f32[le + "0.3.p_swish_alpha"].view(-1), W.f32[le + "0.3.p_swish_beta"].view(-1),
                                  save_stats=save)
        fe = [SimpleNamespace(y=y, m=m, r=r, a=act, cfg=cfg0)]
        for i, (ch, k, st) in enumerate(self.layers[1:], start=1):
            n = le + f"{i}.0.weight"
            xin = fe[-1].a
            b, tin, cinp = xin.shape
            if st > 1:
                pad = int(math.ceil(st / 2))  # nn/utils.py:1089
                tout = (tin + 2 * pad - k) // st + 1
                if not self.fp32 and gemm.strided_conv_ok(tin, cinp, k, st, pad):
                    # input viewed as (B, T/st, st*C): no im2col buffer (85 MB per clip in the large config)
                    col = None
                    y = gemm.strided_conv_nt(xin, W.fwd[n], taps=k, stride=st, pad=pad, out_dtype=self.adt)
                else:
                    col = ops.im2col(xin, k, st, pad, tout)
                    y = self.lin(col, W, n)
                    if not (save and self.keep_im2col):
                        col = None
            else:
                col = None
                y = self.conv(xin, W, n, taps=k, pad=(k - 1) // 2, groups=1)  # padding="same"
            cfg_i = ops.RowLnCfg(ch, 1e-5, act=1)
            act, m, r = ops.rowln_fwd(cfg_i, y, None, W.f32[le + f"{i}.2.1.weight"], W.f32[le + f"{i}.2.1.bias"],
                                      save_stats=save)
            fe.append(SimpleNamespace(y=y, m=m, r=r, a=act, cfg=cfg_i, col=col))
            if not save:
                fe[-2] = None
        clast = self.layers[-1][0]
        cfgp = ops.RowLnCfg(clast, 1e-5)
        lnp, m, r = ops.rowln_fwd(cfgp, fe[-1].a, None, W.f32[ENC + "project_features.1.weight"],
                                  W.f32[ENC + "project_features.1.bias"], save_stats=save or save_proj)
        lf = self.lin(lnp, W, ENC + "project_features.2.weight", bias=W.f32[ENC + "project_features.2.bias"])
        if save:
            c.fe, c.filt, c.x = fe, filt, x
        if save or save_proj:
            c.proj = SimpleNamespace(lnp=lnp, m=m, r=r, cfg=cfgp, a=fe[-1].a)
        return lf

    def _posconv_forward(self, W: _Weights, x: torch.Tensor, save: Optional[list], kept=None) -> torch.Tensor:
        """audio.py:93-113: depth x [grouped Conv1d(k, pad k//2) + bias, LN(no affine), GELU].

        ``kept`` (a MaskIndex, bf16 student path): only the kept positions of the LAST layer's output are ever read
        (base.py:278-280 gathers x_pos at ids_keep), so that layer -- conv, LayerNorm, GELU and in the backward its
        LayerNorm and weight gradient -- runs on the R*Tk kept rows instead of all R*T: the +-k/2 neighbourhood of every
        kept frame is gathered into a (rows, taps, D) operand and the grouped conv becomes one tap-blocked GEMM. Returns
        (R*Tk, D) in kept-row order then, (R, T, D) otherwise."""
        if self.kp % 2 == 0:
            raise NotImplementedError("even positional-conv kernel (SamePad trim)")
        g = self.a.conv_pos_groups
        cfg_l = ops.RowLnCfg(self.D, 1e-5, act=1)
        last = len(self.pos_names) - 1
        for li, n in enumerate(self.pos_names):
            bias = W.f32[n[:-6] + "bias"]
            if kept is not None and li == last:
                rows, t = x.shape[0], x.shape[1]
                nidx = ops.neigh_index(kept.ids_keep, t, self.kp, self.kp // 2)
                xg = ops.row_gather(x.view(rows * t, self.D), nidx, nidx.numel()).view(-1, self.kp, self.D)
                y = gemm.gathered_conv_nt(xg, W.fwd[n], taps=self.kp, groups=g, bias=bias, out_dtype=self.adt)
                act, m, r = ops.rowln_fwd(cfg_l, y, save_stats=save is not None)
                if save is not None:
                    save.append(SimpleNamespace(xg=xg, y=y, m=m, r=r, sparse=True))
                return act
            y = self.conv(x, W, n, taps=self.kp, pad=self.kp // 2, groups=g, bias=bias)
            act, m, r = ops.rowln_fwd(cfg_l, y, save_stats=save is not None)
            if save is not None:
                save.append(SimpleNamespace(x=x, y=y, m=m, r=r, sparse=False))
            x = act
        return x

    def _block_forward(self, W: _Weights, pre: str, x, rows, seq, pos, train: bool, save: Optional[list], site: int):
        """AltBlock.forward, post-LN branch (modules.py:328-337) with AltAttention (:368-410) and timm Mlp."""
        cfg, d = self.cfg, self.D
        f = W.f32
        qkv = self.lin(x, W, pre + "attn.qkv.weight", bias=f[pre + "attn.qkv.bias"])
        s_att, s1, s2 = self._seed(site), self._seed(site + 1), self._seed(site + 2)
        p_att = cfg.attention_dropout if train else 0.0
        ao, lse = ops.attn_fwd(qkv.view(rows, seq, 3 * d), rows, seq, self.H, pos=pos, slopes=self.slopes,
                               alibi_scale=W.alibi, drop_p=p_att, seed=s_att, need_lse=save is not None,
                               skip_far_keys=self.alibi_locality and pos is None)
        pr = self.lin(ao.view(rows * seq, d), W, pre + "attn.proj.weight", bias=f[pre + "attn.proj.bias"])
        c1 = ops.RowLnCfg(d, cfg.norm_eps, drop_b=cfg.encoder_dropout)
        x1, m1, r1 = ops.rowln_fwd(c1, x, pr, f[pre + "norm1.weight"], f[pre + "norm1.bias"], seed_b=s1,
                                   training=train, save_stats=save is not None)
        u = torch.empty(rows * seq, self.hidden, device=x.device, dtype=self.adt) if save is not None else None
        h = self.lin(x1, W, pre + "mlp.fc1.weight", bias=f[pre + "mlp.fc1.bias"], act=1, preact=u)
        p_mlp = cfg.activation_dropout if train else 0.0  # timm Mlp drop1 (= mlp_drop, modules.py:312-317)
        s_mlp = self._seed(site + 3)
        if p_mlp > 0:
            h = ops.row_gather(h, self._arange(rows * seq), rows * seq, drop_p=p_mlp, drop_seed=s_mlp)
        t = self.lin(h, W, pre + "mlp.fc2.weight", bias=f[pre + "mlp.fc2.bias"])
        c2 = ops.RowLnCfg(d, cfg.norm_eps, drop_b=cfg.post_mlp_drop)
        x2, m2, r2 = ops.rowln_fwd(c2, x1, t, f[pre + "norm2.weight"], f[pre + "norm2.bias"], seed_b=s2,
                                   training=train, save_stats=save is not None)
        if save is not None:
            save.append(SimpleNamespace(pre=pre, x=x, qkv=qkv, ao=ao, lse=lse, pr=pr, x1=x1, m1=m1, r1=r1, u=u, h=h,
                                        t=t, m2=m2, r2=r2, s_att=s_att, s1=s1, s2=s2, p_att=p_att, c1=c1, c2=c2,
                                        p_mlp=p_mlp, s_mlp=s_mlp))
        return x2, t

    def _encoder_forward(self, W: _Weights, x, rows, seq, pos, train, save: Optional[list], targets: Optional[list],
                         c: Optional[SimpleNamespace]):
        """BlockEncoder (modules.py:83-108: LN -> dropout -> prenet blocks) followed by the main blocks."""
        cfg = self.cfg
        cn = ops.RowLnCfg(self.D, cfg.norm_eps, drop_out=self.a.prenet_dropout)
        s_pre = self._seed(1)
        xn, m, r = ops.rowln_fwd(cn, x, None, W.f32[ENC + "context_encoder.norm.weight"],
                                 W.f32[ENC + "context_encoder.norm.bias"], seed_out=s_pre, training=train,
                                 save_stats=save is not None)
        if save is not None:
            c.prenorm = SimpleNamespace(x=x, m=m, r=r, cfg=cn, seed=s_pre)
        x = xn
        for j, pre in enumerate(self.block_prefixes):
            x, t = self._block_forward(W, pre, x, rows, seq, pos, train, save, 16 + 4 * j)
            if targets is not None and j >= self.a.prenet_depth:
                targets.append(t.view(rows, seq, self.D))
        return x

    def _decoder_forward(self, W: _Weights, x: torch.Tensor, save: Optional[list]) -> torch.Tensor:
        """Decoder1d.forward (modules.py:179-192) on the group-padded layout; residual rule of :124-134."""
        dc = self.dec
        if dc.decoder_kernel % 2 == 0:
            raise NotImplementedError("even decoder kernel (SamePad trim)")
        cfg_l = ops.RowLnCfg(self.dec_wp, 1e-5, act=1, group_width=self.dec_gw, group_real=self.dec_ng)
        compact = self.dec_gw != GP
        for l in range(dc.decoder_layers):
            n = ENC + f"decoder.blocks.{l}.0.weight"
            y = self.conv(x, W, n, taps=dc.decoder_kernel, pad=dc.decoder_kernel // 2, groups=dc.decoder_groups,
                          bias=W.f32[ENC + f"decoder.blocks.{l}.0.bias"], x_real=0 if l == 0 else self.dec_ng,
                          ng_out=self.dec_ng if compact else None)
            res = x if l > 0 else None  # layer 0 changes the channel count: no residual
            act, m, r = ops.rowln_fwd(cfg_l, y, post=res, save_stats=save is not None)
            if save is not None:
                save.append(SimpleNamespace(x=x, y=y, m=m, r=r, res=res is not None, cfg=cfg_l))
            x = act
        return self.lin(x, W, ENC + "decoder.proj.weight", bias=W.f32[ENC + "decoder.proj.bias"]), x

    # ------------------------------------------------------------------------------------ forward
    def forward(self, source: torch.Tensor, ids=None, num_updates: int = 0, *, mask: Optional[np.ndarray] = None,
                training: bool = True, need_grad: bool = True, taps: Optional[dict] = None,
                fuse_loss_grad: bool = False) -> Dict[str, object]:
        """One pretraining forward. Returns loss_sum (device double scalar), sample_size, the logging statistics
        and keeps what the backward needs in ``self.ctx``. ``fuse_loss_grad`` (bf16): the loss kernel also writes
        d loss_sum / d pred over the prediction in the same pass; the following ``backward`` must then be called
        without an upstream gradient scalar (the step driver's case)."""
        cfg, d, M = self.cfg, self.D, self.M
        L.require_device(source)
        if not (self.has_teacher and self.has_decoder):
            raise RuntimeError("pretraining forward after remove_pretraining_modules()")
        self._refresh_student()
        self._refresh_teacher()
        self.step_counter += 1
        c = SimpleNamespace()
        save = need_grad
        x = source.to(torch.float32).contiguous()
        B, N = x.shape
        if training and cfg.source_mixup >= 0 and cfg.mixup_prob > 0:
            x = self._mixup(x)
            if taps is not None:
                taps["mixed_source"] = x
        lf = self._fe_forward(x, c, save)
        T = lf.shape[1]
        if taps is not None:
            taps["local_features"] = lf
            if save:
                for i, s in enumerate(c.fe):
                    taps[f"fe_layer{i}"] = s.a

        # ---- masks (host integer work, bit exact) and the device index maps
        if mask is None:
            idl = None if ids is None else [int(v) for v in (ids.tolist() if hasattr(ids, "tolist") else ids)]
            mask = self._prefetcher(B, T).get(num_updates, idl) if idl is not None else \
                masking.pretrain_mask(update=num_updates, ids=None, **self._mask_static(B, T))
        R = B * M
        assert mask.shape == (R, T)
        tk = int(T - mask[0].sum())
        mask_u8 = ops.h2d_async(torch.from_numpy(np.ascontiguousarray(mask).view(np.uint8)), self.device)
        mi = ops.mask_index(mask_u8, tk, M)
        c.mi, c.B, c.T, c.tk, c.R = mi, B, T, tk, R
        n_masked = int(R * (T - tk))

        # ---- student: clone + zero-mask, positional conv, keep unmasked, encoder, decoder
        lf2 = lf.view(B * T, d)
        x_masked = ops.row_gather(lf2, mi.clone_src, R * T, out_shape=(R, T, d))
        c.pos = [] if save else None
        sparse_last = self.sparse_last_posconv and not self.fp32 and (d // self.a.conv_pos_groups) % 64 == 0
        x_pos = self._posconv_forward(self.WS, x_masked, c.pos, kept=mi if sparse_last else None)
        x_unm = ops.row_gather(lf2, mi.keep_src_x, R * tk, out_shape=(R * tk, d))
        if sparse_last:  # x_pos is already (R*Tk, D) in kept-row order
            xs = self._add(x_pos, x_unm)
        else:
            xs = ops.row_gather(x_pos.view(R * T, d), mi.keep_src_clone, R * tk, add=x_unm, out_shape=(R * tk, d))
        del x_pos, x_unm, x_masked
        c.blocks = [] if save else None
        xs = self._encoder_forward(self.WS, xs, R, tk, mi.ids_keep, training, c.blocks, None, c)
        if taps is not None:
            taps["student_out"] = ops.row_gather(xs, mi.restore_src, R * T, out_shape=(R, T, d))
            if save:
                taps["student_prenet"] = ops.row_gather(c.blocks[self.a.prenet_depth].x, mi.restore_src, R * T,
                                                        out_shape=(R, T, d))
        s_dec, s_noise = self._seed(2), self._seed(3)
        p_dec = self.dec.input_dropout if training else 0.0
        noise = self.a.mask_noise_std
        dec_in = ops.row_gather(xs, mi.restore_src, R * T, fill_std=noise, fill_seed=s_noise, drop_p=p_dec,
                                drop_seed=s_dec, drop_by_src=True, out_shape=(R, T, d))
        c.dec = [] if save else None
        pred, dec_last = self._decoder_forward(self.WS, dec_in, c.dec)
        c.pred, c.dec_last, c.s_dec, c.p_dec = pred, dec_last, s_dec, p_dec
        if taps is not None:
            taps["decoder_out"] = pred

        # ---- teacher (no grad, eval mode): full-length positional conv + encoder, top-K FFN targets
        y_pos = self._posconv_forward(self.WT, lf, None)
        ty = ops.row_gather(lf2, self._arange(B * T), B * T, add=y_pos.view(B * T, d), out_shape=(B * T, d))
        del y_pos
        targets: List[torch.Tensor] = []
        self._encoder_forward(self.WT, ty, B, T, None, False, None, targets, None)
        y = ops.make_targets(targets[-cfg.average_top_k_layers:], 1e-5)
        del targets
        if taps is not None:
            taps["targets"] = y

        # ---- masked regression loss + collapse statistics
        scale = cfg.loss_scale if cfg.loss_scale is not None else 1.0 / math.sqrt(d)
        scale = float(scale) * float(cfg.d2v_loss)
        c.loss_grad_fused = bool(fuse_loss_grad and save and not self.fp32 and taps is None and d % 256 == 0)
        if c.loss_grad_fused:
            loss_sum, stats = ops.d2v_loss_fused(pred.view(R, T, d), y, mask_u8, M, scale, 2.0 * scale)
        else:
            loss_sum, stats = ops.d2v_loss_fwd(pred.view(R, T, d), y, mask_u8, M, scale)
        c.y, c.mask_u8, c.scale = y, mask_u8, scale
        c.lf = lf
        self.ctx = c if save else None
        return {"loss_sum": loss_sum, "colstats": stats, "sample_size": n_masked, "masked_pct": 1.0 - tk / T,
                "mask": mask, "T": T, "tk": tk}

    # ------------------------------------------------------------------------------------ features_only path
    def extract_features(self, source: torch.Tensor, ids=None, num_updates: int = 0, *, mask: bool = False,
                         precomputed_mask: Optional[np.ndarray] = None, training: bool = False,
                         need_grad: bool = False) -> Dict[str, object]:
        """Data2VecMultiModel.forward(features_only=True) (nn/data2vec2.py:632-728; extract_features :1112-1123):
        the STUDENT on the full-length sequence, clone_batch 1, masked rows kept in place. Returns ``x`` (B, T, D),
        ``layer_results`` (FFN outputs of the main blocks, what the finetune head averages, nn/wav2vec2.py:446-462)
        and the mask. Unmasked eval mode is the README inference contract (README.md:69-121)."""
        L.require_device(source)
        if mask or training or need_grad:
            raise NotImplementedError("features_only with masking / training mode / gradients: use FinetuneEngine")
        self._refresh_student()
        d = self.D
        x = source.to(torch.float32).contiguous()
        B = x.shape[0]
        lf = self._fe_forward(x, SimpleNamespace(), False)
        T = lf.shape[1]
        x_pos = self._posconv_forward(self.WS, lf, None)
        xs = self._add(x_pos.view(B * T, d), lf.view(B * T, d))
        del x_pos
        layer_results: List[torch.Tensor] = []
        xs = self._encoder_forward(self.WS, xs, B, T, None, False, None, layer_results, None)
        return {"x": xs.view(B, T, d), "layer_results": layer_results, "mask": None, "local_features": lf}

    def _mask_static(self, B: int, T: int) -> dict:
        return dict(seed=self.cfg.seed, batch=B, frames=T, clone_batch=self.M, mask_prob=self.a.mask_prob,
                    mask_length=self.a.mask_length, mask_dropout=self.a.mask_dropout, add_masks=self.a.add_masks)

    def _prefetcher(self, B: int, T: int) -> masking.MaskPrefetcher:
        key = (B, T)
        pf = getattr(self, "_pf", None)
        if pf is None or pf[0] != key:
            if pf is not None:
                pf[1].close()
            pf = self._pf = (key, masking.MaskPrefetcher(**self._mask_static(B, T)))
        return pf[1]

    def frames_for(self, n_samples: int) -> int:
        """Frames the feature extractor produces for ``n_samples`` input samples."""
        t = n_samples
        for (_c, k, st) in self.layers[1:]:
            if st > 1:
                t = (t + 2 * int(math.ceil(st / 2)) - k) // st + 1
        return t

    def prefetch_mask(self, num_updates: int, ids, batch: int, n_samples: int) -> None:
        """Start computing the masks of a FUTURE forward(ids, num_updates) on the worker thread: they depend
        only on (seed, update, ids), so the host numpy work overlaps the GPU step before it."""
        self._prefetcher(batch, self.frames_for(n_samples)).announce(num_updates, [int(v) for v in ids])

    def _arange(self, n: int) -> torch.Tensor:
        t = getattr(self, "_arange_buf", None)
        if t is None or t.numel() < n:
            t = self._arange_buf = torch.arange(n, device=self.device, dtype=torch.int32)
        return t[:n]

    # ------------------------------------------------------------------------------------ backward
    def _block_backward(self, W: _Weights, s: SimpleNamespace, dx2, rows, seq, pos, train: bool, extra_dt=None):
        """``extra_dt``: additional gradient of the block's FFN output ``t`` (the finetune head averages the FFN
        outputs of the top-k blocks, nn/wav2vec2.py:446-462); ``dx2`` None: the block output itself is unused."""
        pre, G, f = s.pre, self.G, W.f32
        if dx2 is None:
            assert extra_dt is not None
            dz2, dt = None, extra_dt
            ops.colsum(dt, G(pre + "mlp.fc2.bias"))
        else:
            dz2, dt = ops.rowln_bwd(s.c2, dx2, s.x1, s.t, f[pre + "norm2.weight"], f[pre + "norm2.bias"], None, None,
                                    s.m2, s.r2, seed_b=s.s2, training=train, dgamma=G(pre + "norm2.weight"),
                                    dbeta=G(pre + "norm2.bias"),
                                    dbias_b=G(pre + "mlp.fc2.bias") if extra_dt is None else None)
            if extra_dt is not None:
                dt = self._add(dt, extra_dt)
                ops.colsum(dt, G(pre + "mlp.fc2.bias"))
        self.wgrad(dt, s.h, G(pre + "mlp.fc2.weight"))
        p_mlp = getattr(s, "p_mlp", 0.0)
        if not self.fp32 and p_mlp == 0 and gemm.pair_shape_ok(dt.numel() // dt.shape[-1], s.u.shape[-1], dt.shape[-1],
                                                               dt.dtype):
            # GELU backward and the fc1 bias gradient ride in the epilogue of the fc2 data-gradient product
            du = self.lin(dt, W, pre + "mlp.fc2.weight", dgrad=True, dgelu_u=s.u.view(*dt.shape[:-1], -1),
                          colsum=G(pre + "mlp.fc1.bias"))
        else:
            dh = self.lin(dt, W, pre + "mlp.fc2.weight", dgrad=True)
            if p_mlp > 0:
                dh = ops.row_gather(dh, self._arange(rows * seq), rows * seq, drop_p=s.p_mlp, drop_seed=s.s_mlp)
            du = ops.dgelu_mul(dh, s.u, colsum=G(pre + "mlp.fc1.bias"))
        self.wgrad(du, s.x1, G(pre + "mlp.fc1.weight"))
        dx1 = self.lin(du, W, pre + "mlp.fc1.weight", dgrad=True, residual=dz2)
        del du, dz2, dt
        dz1, dpr = ops.rowln_bwd(s.c1, dx1, s.x, s.pr, f[pre + "norm1.weight"], f[pre + "norm1.bias"], None, None,
                                 s.m1, s.r1, seed_b=s.s1, training=train, dgamma=G(pre + "norm1.weight"),
                                 dbeta=G(pre + "norm1.bias"), dbias_b=G(pre + "attn.proj.bias"))
        self.wgrad(dpr, s.ao.view(rows * seq, self.D), G(pre + "attn.proj.weight"))
        dao = self.lin(dpr, W, pre + "attn.proj.weight", dgrad=True)
        dal = G(ENC + "alibi_scale").view(-1) if self.a.learned_alibi_scale else None
        dqkv = ops.attn_bwd(dao.view(rows, seq, self.D), s.qkv.view(rows, seq, 3 * self.D), s.ao, s.lse, rows, seq,
                            self.H, pos=pos, slopes=self.slopes, alibi_scale=W.alibi, dalibi_scale=dal,
                            drop_p=s.p_att, seed=s.s_att,
                            dqkv_colsum=G(pre + "attn.qkv.bias")).view(rows * seq, 3 * self.D)
        self.wgrad(dqkv, s.x, G(pre + "attn.qkv.weight"))
        return self.lin(dqkv, W, pre + "attn.qkv.weight", dgrad=True, residual=dz1)

    def backward(self, grad_scale: Optional[torch.Tensor] = None, training: bool = True, block_done=None) -> None:
        """Accumulate d(loss_sum * grad_scale)/d(parameters) into the flat gradient buffer. ``block_done(i)``
        is called after the gradients of encoder block i (reverse order) are complete (DDP bucket hook)."""
        c = self.ctx
        if c is None:
            raise RuntimeError("backward() without a preceding forward(need_grad=True)")
        W, G, d = self.WS, self.G, self.D
        mi, B, T, tk, R, M = c.mi, c.B, c.T, c.tk, c.R, self.M
        dc = self.dec
        if c.loss_grad_fused:
            if grad_scale is not None:
                raise RuntimeError("forward(fuse_loss_grad=True) already wrote the gradient for an upstream scalar of 1")
            dpred = c.pred.view(R * T, d)  # overwritten in place by the fused loss kernel
        else:
            dpred = ops.d2v_loss_bwd(c.pred.view(R, T, d), c.y, c.mask_u8, M, c.scale, grad_scale).view(R * T, d)
        # decoder projection
        n = ENC + "decoder.proj.weight"
        ops.colsum(dpred, G(ENC + "decoder.proj.bias"))
        self.wgrad(dpred, c.dec_last.view(R * T, self.dec_wp), self.gpacked[n + "|F"])
        dx = self.lin(dpred, W, n, dgrad=True).view(R, T, self.dec_wp)
        del dpred
        for l in reversed(range(dc.decoder_layers)):
            s = c.dec[l]
            n = ENC + f"decoder.blocks.{l}.0.weight"
            dy, _ = ops.rowln_bwd(s.cfg, dx, s.y, None, None, None, None, None, s.m, s.r)
            ops.colsum(dy.view(R * T, self.dec_wp), self.gpacked[ENC + f"decoder.blocks.{l}.0.bias|B"])
            self.conv_wgrad(dy, s.x, self.gpacked[n + "|F"], taps=dc.decoder_kernel, pad=dc.decoder_kernel // 2,
                            groups=dc.decoder_groups)
            dxin = self.conv(dy, W, n, taps=dc.decoder_kernel, pad=dc.decoder_kernel - 1 - dc.decoder_kernel // 2,
                             groups=dc.decoder_groups, dgrad=True, x_real=self.dec_ng,
                             ng_out=self.dec_ng if (self.dec_gw != GP and l > 0) else None)
            if s.res:  # y = act(...) + x: the residual passes dx straight through
                dxin = self._add(dxin, dx)
            dx = dxin
            c.dec[l] = None
        # decoder input: kept tokens came from xs (with input dropout), mask tokens were noise
        dxs = ops.row_gather(dx.view(R * T, d), mi.keep_src_clone, R * tk, drop_p=c.p_dec, drop_seed=c.s_dec,
                             out_shape=(R * tk, d))
        del dx
        nb = len(self.block_prefixes)
        for j in reversed(range(nb)):
            dxs = self._block_backward(W, c.blocks[j], dxs, R, tk, mi.ids_keep, training)
            c.blocks[j] = None
            if block_done is not None:
                block_done(j)
        s = c.prenorm
        dxs, _ = ops.rowln_bwd(s.cfg, dxs, s.x, None, W.f32[ENC + "context_encoder.norm.weight"],
                               W.f32[ENC + "context_encoder.norm.bias"], None, None, s.m, s.r, seed_out=s.seed,
                               training=training, dgamma=G(ENC + "context_encoder.norm.weight"),
                               dbeta=G(ENC + "context_encoder.norm.bias"))
        # xs0 = x_unmasked + x_pos[ids_keep]
        g = self.a.conv_pos_groups
        cfg_l = ops.RowLnCfg(d, 1e-5, act=1)
        dpos = None
        for li in reversed(range(len(self.pos_names))):
            s = c.pos[li]
            n = self.pos_names[li]
            if s.sparse:  # last layer, evaluated on the kept rows only (see _posconv_forward)
                dyk, _ = ops.rowln_bwd(cfg_l, dxs, s.y, None, None, None, None, None, s.m, s.r)
                ops.colsum(dyk, G(n[:-6] + "bias"))
                gemm.gathered_conv_wgrad_tn(dyk, s.xg, self.gpacked[n + "|F"], taps=self.kp, groups=g)
                dy = ops.row_gather(dyk, mi.restore_src, R * T, out_shape=(R, T, d))  # zeros at the masked frames
                dpos = self.conv(dy, W, n, taps=self.kp, pad=self.kp - 1 - self.kp // 2, groups=g, dgrad=True)
                c.pos[li] = None
                continue
            if dpos is None:
                dpos = ops.row_gather(dxs, mi.restore_src, R * T, out_shape=(R, T, d))
            dy, _ = ops.rowln_bwd(cfg_l, dpos, s.y, None, None, None, None, None, s.m, s.r)
            ops.colsum(dy.view(R * T, d), G(n[:-6] + "bias"))
            self.conv_wgrad(dy, s.x, self.gpacked[n + "|F"], taps=self.kp, pad=self.kp // 2, groups=g)
            dpos = self.conv(dy, W, n, taps=self.kp, pad=self.kp - 1 - self.kp // 2, groups=g, dgrad=True)
            c.pos[li] = None
        dlf = ops.clone_sum_bwd(dpos, dxs, mi.restore_src, B, T, M, d)
        del dpos, dxs
        self._fe_backward(c, dlf.view(B * T, d))
        self._unpack_grads()
        self.ctx = None

    def _add(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        n = a.numel()
        return ops.row_gather(a.view(-1, a.shape[-1]), self._arange(n // a.shape[-1]), n // a.shape[-1],
                              add=b.view(-1, a.shape[-1]), out_shape=a.shape)

    def _proj_backward(self, c: SimpleNamespace, dlf: torch.Tensor) -> torch.Tensor:
        """Backward of project_features (LayerNorm(C) + Linear(C -> D), audio.py:83-88): parameter gradients and the
        gradient of the conv extractor's output."""
        W, G = self.WS, self.G
        n = ENC + "project_features.2.weight"
        ops.colsum(dlf, G(ENC + "project_features.2.bias"))
        self.wgrad(dlf, c.proj.lnp, G(n))
        dl = self.lin(dlf, W, n, dgrad=True)
        da, _ = ops.rowln_bwd(c.proj.cfg, dl.view(c.proj.a.shape), c.proj.a, None,
                              W.f32[ENC + "project_features.1.weight"], W.f32[ENC + "project_features.1.bias"], None,
                              None, c.proj.m, c.proj.r, dgamma=G(ENC + "project_features.1.weight"),
                              dbeta=G(ENC + "project_features.1.bias"))
        return da

    def _fe_backward(self, c: SimpleNamespace, dlf: torch.Tensor) -> None:
        W, G = self.WS, self.G
        le = ENC + "local_encoder.conv_layers."
        fe = c.fe
        da = self._proj_backward(c, dlf)
        for i in reversed(range(1, len(self.layers))):
            ch, k, st = self.layers[i]
            s, xin = fe[i], fe[i - 1].a
            n = le + f"{i}.0.weight"
            dy, _ = ops.rowln_bwd(s.cfg, da, s.y, None, W.f32[le + f"{i}.2.1.weight"], W.f32[le + f"{i}.2.1.bias"],
                                  None, None, s.m, s.r, dgamma=G(le + f"{i}.2.1.weight"),
                                  dbeta=G(le + f"{i}.2.1.bias"))
            b, tin, cinp = xin.shape
            if st > 1:
                pad = int(math.ceil(st / 2))
                tout = dy.shape[1]
                if not self.fp32 and gemm.strided_conv_ok(tin, cinp, k, st, pad):
                    gemm.strided_conv_wgrad_tn(dy, xin, self.gpacked[n + "|F"], taps=k, stride=st, pad=pad)
                    da = gemm.strided_conv_dgrad(dy, W.dgrad_s[n], c=cinp, taps_per_block=-(-k // st), stride=st, pad=pad,
                                                 out_dtype=self.adt)
                else:
                    col = s.col if s.col is not None else ops.im2col(xin, k, st, pad, tout)
                    if self.fp32:
                        self.wgrad(dy, col, self.gpacked[n + "|F"])
                    else:  # same transposed (k*C, N) layout as the strided path writes
                        gemm.gemm_tn(col.reshape(-1, col.shape[-1]), dy.reshape(-1, dy.shape[-1]), self.gpacked[n + "|F"])
                    del col
                    s.col = None
                    dcol = self.lin(dy, W, n, dgrad=True)
                    da = ops.col2im(dcol.view(b, tout, k * cinp), k, st, pad, tin)
                    del dcol
            else:
                self.conv_wgrad(dy, xin, self.gpacked[n + "|F"], taps=k, pad=(k - 1) // 2, groups=1)
                da = self.conv(dy, W, n, taps=k, pad=k - 1 - (k - 1) // 2, groups=1, dgrad=True)
            fe[i] = None
        s = fe[0]
        c0, k0, _ = self.layers[0]
        dy0, _ = ops.rowln_bwd(s.cfg, da, s.y, None, W.f32[le + "0.2.1.weight"], W.f32[le + "0.2.1.bias"],
                               W.f32[le + "0.3.p_swish_alpha"].view(-1), W.f32[le + "0.3.p_swish_beta"].view(-1),
                               s.m, s.r, dgamma=G(le + "0.2.1.weight"), dbeta=G(le + "0.2.1.bias"),
                               dact_alpha=G(le + "0.3.p_swish_alpha").view(-1),
                               dact_beta=G(le + "0.3.p_swish_beta").view(-1))
        dfilt = ops.sinc_conv_wgrad(c.x, dy0, k0)
        ops.sinc_filters_bwd(W.f32[le + "0.0.low_hz_"].view(-1), W.f32[le + "0.0.band_hz_"].view(-1), self.sinc_n,
                             self.sinc_window, k0, self.min_low_hz, self.min_band_hz, float(self.a.sample_rate), dfilt,
                             G(le + "0.0.low_hz_").view(-1), G(le + "0.0.band_hz_").view(-1))

    def _unpack_grads(self) -> None:
        """Packed-layout weight gradients -> checkpoint-layout views of the flat gradient buffer."""
        if not self.fp32:
            if self._unpack_table is None:
                t = self._unpack_table = ops.RelayoutTable(self.device)
                for key, buf in self.gpacked.items():
                    name, _ = key.split("|")
                    pk = self.sp[key]
                    strides = pk.gt_strides if key in self.gt_keys else pk.out_strides
                    # G += packed, and the packed accumulator is cleared for the next backward in the same pass
                    t.add(buf, self.G(name), pk.dims, strides, 0, pk.in_strides, pk.in_off, accumulate=True, zero_src=True)
            self._unpack_table.run()
            return
        for key, buf in self.gpacked.items():
            name, _ = key.split("|")
            P.unpack_grad(self.sp[key], buf, self.G(name), transposed=key in self.gt_keys)  # G += packed
            buf.zero_()

    def zero_grad(self) -> None:
        self.S.grad.zero_()

    # ------------------------------------------------------------------------------------ EMA
    def ema_step(self, num_updates: int) -> float:
        """set_num_updates (nn/data2vec2.py:386-410): anneal the decay, then fairseq EMAModule.step on the
        shared parameters (fp32 shadow) and refresh the teacher's bf16 copy -- one fused launch."""
        decay = annealed_decay(self.cfg, num_updates)
        if decay < 1 and self.has_teacher:
            n = self.E.total
            ops.ema_step(self.S.data[:n], self.E.data, None if self.fp32 else self.T16, decay)
            self._teacher_dirty = True
            self._t16_valid = not self.fp32
        return decay

    # ------------------------------------------------------------------------------------ logging statistics
    @staticmethod
    def variances(stats: torch.Tensor, n: int):
        """compute_var (nn/data2vec2.py:1095-1110) from the fused column sums [sum x, sum x^2, sum y, sum y^2]."""
        sx, sxx, sy, syy = stats
        var_x = (sxx - sx * sx / n) / (n - 1)
        var_y = (syy - sy * sy / n) / (n - 1)
        return torch.sqrt(var_x + 1e-6).mean(), torch.sqrt(var_y + 1e-6).mean()
