"""Finetune step (BASELINE configs[3]) as an explicit schedule of the sm_100a kernels.

Replaces, under /root/reference:
  nn/wav2vec2.py:362-482      Wav2VecEncoderModOut.forward: BC mixup of source (+ targets), extract_features in
                              train mode with masking, mean of the top-k FFN outputs, Linear(D -> classes)
  nn/data2vec2.py:632-728     Data2VecMultiModel.forward(features_only=True): clone_batch 1, masked rows stay in
                              place (remove_masked=False), layerdrop over the main blocks, no decoder, no teacher
  nn/modalities/base.py:457-484 apply_mask with encoder_zero_mask=False (mask tokens ~ N(0, mask_noise_std)) and
                              channel masking; :194-213 local_features under local_grad_mult = 0 (no gradient)
  nn/criterions.py:231-277    FinetuneCrossEntropyCriterion.forward, focal branch + accuracy / confusion counters
and their autograd backward: full-length (T = 2000) student blocks with gradients -- the tiled attention backward of
csrc/attention_bwd.cu -- the positional-conv weight gradients, the head. ``freeze_finetune_updates`` (frozen phase:
only the head trains, the encoder runs without saving activations) follows nn/wav2vec2.py:437-441.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np
import torch

from . import lib as L
from . import masking, ops
from .config import Data2VecMultiConfig, Wav2Vec2CcasFinetuneConfig, finetune_overrides
from .engine import ENC, PretrainEngine


class FinetuneEngine:
    def __init__(self, model_cfg: Data2VecMultiConfig, ft_cfg: Wav2Vec2CcasFinetuneConfig, num_classes: int,
                 device="cuda", precision: str = "bf16", init: Optional[Dict[str, torch.Tensor]] = None,
                 head_init: Optional[Dict[str, torch.Tensor]] = None, init_seed: int = 0, rng_seed: int = 0,
                 metric_threshold: float = 0.25):
        self.ft = ft_cfg
        self.metric_threshold = float(metric_threshold)  # criterion.metric_threshold (nn/criterions.py:95-98)
        cfg = finetune_overrides(model_cfg, ft_cfg)  # nn/wav2vec2.py:95-130 arg_overrides
        if ft_cfg.source_mixup >= 0 and ft_cfg.mixup_prob > 0 and \
                (not ft_cfg.same_mixup or ft_cfg.mixup_prob < 1 or ft_cfg.gain_mode != "A_weighting"):
            raise NotImplementedError("finetune mixup variant other than same_mixup / mixup_prob=1 / A_weighting")
        if ft_cfg.final_dropout:
            raise NotImplementedError("final_dropout > 0 (0 in the shipped finetune recipes)")
        cfg.source_mixup, cfg.mixing_window_length = ft_cfg.source_mixup, ft_cfg.mixing_window_length
        cfg.mixup_prob, cfg.same_mixup, cfg.gain_mode = max(ft_cfg.mixup_prob, 0.0), True, "A_weighting"
        if ft_cfg.source_mixup >= 0 and ft_cfg.mixup_prob <= 0:
            cfg.source_mixup = -1.0
        self.core = PretrainEngine(cfg, device, precision=precision, init=init, init_seed=init_seed, rng_seed=rng_seed,
                                   finetune=True)
        self.core.drop_teacher()   # model.remove_pretraining_modules (nn/wav2vec2.py:187)
        self.core.drop_decoder()
        self.cfg = cfg
        self.device = self.core.device
        self.C = int(num_classes)
        self.D = self.core.D
        self.top_k = int(ft_cfg.average_top_k_layers)
        # proj = Linear(D, classes): xavier_uniform weight, zero bias (nn/wav2vec2.py:210-211, :1144-1149)
        n_head = self.C * self.D + (self.C + 3) // 4 * 4
        self.head = torch.zeros(n_head, device=self.device, dtype=torch.float32)
        self.head_grad = torch.zeros_like(self.head)
        if head_init is not None:
            self.head_w.copy_(head_init["proj.weight"].to(self.device, torch.float32))
            self.head_b.copy_(head_init["proj.bias"].to(self.device, torch.float32))
        else:
            gen = torch.Generator().manual_seed(init_seed + 17)
            bound = math.sqrt(6.0 / (self.D + self.C))
            self.head_w.copy_(((torch.rand(self.C, self.D, generator=gen) * 2 - 1) * bound).to(self.device))
        self.num_updates = 0
        self.ctx = None

    # ------------------------------------------------------------------------------------ head parameter views
    @property
    def head_w(self) -> torch.Tensor:
        return self.head[: self.C * self.D].view(self.C, self.D)

    @property
    def head_b(self) -> torch.Tensor:
        return self.head[self.C * self.D: self.C * self.D + self.C]

    @property
    def head_gw(self) -> torch.Tensor:
        return self.head_grad[: self.C * self.D].view(self.C, self.D)

    @property
    def head_gb(self) -> torch.Tensor:
        return self.head_grad[self.C * self.D: self.C * self.D + self.C]

    def zero_grad(self) -> None:
        self.core.zero_grad()
        self.head_grad.zero_()

    @property
    def encoder_trainable(self) -> bool:
        """nn/wav2vec2.py:437: ft = freeze_finetune_updates <= num_updates."""
        return self.ft.freeze_finetune_updates <= self.num_updates

    # ------------------------------------------------------------------------------------ forward
    def forward(self, source: torch.Tensor, target: Optional[torch.Tensor] = None, *, training: bool = True,
                time_mask: Optional[np.ndarray] = None, channel_mask: Optional[np.ndarray] = None,
                need_grad: Optional[bool] = None, want_unreduced: bool = False) -> Dict[str, object]:
        """One finetune forward. ``target`` (B, T, classes) multi-hot / soft labels. Returns the logits
        (``encoder_out``, B x T x C fp32), and when a target is given the summed focal loss (device double scalar)
        and the accuracy / confusion counters. ``time_mask`` (B, T) / ``channel_mask`` (B, D) bool override the
        random masks (tests); ``need_grad`` defaults to ``training``."""
        e, cfg, ft, d = self.core, self.cfg, self.ft, self.D
        L.require_device(source)
        e._refresh_student()
        e.step_counter += 1
        need_grad = training if need_grad is None else need_grad
        enc_grad = need_grad and self.encoder_trainable
        c = SimpleNamespace()
        x = source.to(torch.float32).contiguous()
        B = x.shape[0]
        mix = {}
        if training and cfg.source_mixup >= 0 and cfg.mixup_prob > 0:
            x = e._mixup(x, mix)
        # feature_grad_mult 0: the conv extractor is frozen, project_features still trains (base.py:205-207)
        lf = e._fe_forward(x, c, False, save_proj=enc_grad)
        T = lf.shape[1]
        a = e.a
        tmask = cmask = None
        xin = lf.view(B * T, d)
        if training and ft.apply_mask:
            # nn/modalities/base.py:370-425 (clone_batch 1, no mask seeds: OS entropy) + apply_mask :457-484
            if time_mask is None and a.mask_prob > 0:
                time_mask = masking.pretrain_mask(seed=cfg.seed, update=self.num_updates, ids=None, batch=B, frames=T,
                                                  clone_batch=1, mask_prob=a.mask_prob, mask_length=a.mask_length,
                                                  mask_dropout=a.mask_dropout, add_masks=a.add_masks)
            if time_mask is not None:
                tmask = np.ascontiguousarray(time_mask).astype(bool)
                assert tmask.shape == (B, T)
                if a.encoder_zero_mask:
                    raise NotImplementedError("features_only masking with encoder_zero_mask=True")
                idx = np.where(tmask.reshape(-1), -1, np.arange(B * T)).astype(np.int32)
                idx_d = ops.h2d_async(torch.from_numpy(idx), self.device)
                xin = ops.row_gather(xin, idx_d, B * T, fill_std=a.mask_noise_std, fill_seed=e._seed(5))
                c.keep_idx = idx_d
            if channel_mask is None and a.mask_channel_prob > 0:
                channel_mask = masking.compute_mask_indices(B, d, a.mask_channel_prob, a.mask_channel_length, seed=None,
                                                            epoch=None, indices=None)
            if channel_mask is not None:
                cmask = np.ascontiguousarray(channel_mask).astype(bool)
                assert cmask.shape == (B, d)
                cm_d = ops.h2d_async(torch.from_numpy(cmask.view(np.uint8)), self.device)
                if xin.data_ptr() == lf.data_ptr():
                    xin = xin.clone()
                ops.channel_mask_(xin, cm_d, T)
                c.cm_d = cm_d
        # positional encoder on the (masked) full-length sequence, then x + x_pos (base.py:268-283)
        c.pos = [] if enc_grad else None
        x_pos = e._posconv_forward(e.WS, xin.view(B, T, d), c.pos)
        xs = e._add(x_pos.view(B * T, d), xin)
        del x_pos
        # prenet (BlockEncoder) + main blocks with layerdrop (nn/data2vec2.py:652-676; prenet: modules.py:96-104)
        c.blocks = [] if enc_grad else None
        layer_results: List[torch.Tensor] = []
        ld_main = cfg.layerdrop if training else 0.0
        ld_pre = a.prenet_layerdrop if training else 0.0
        cn = ops.RowLnCfg(d, cfg.norm_eps, drop_out=a.prenet_dropout)
        s_pre = e._seed(1)
        xn, m, r = ops.rowln_fwd(cn, xs, None, e.WS.f32[ENC + "context_encoder.norm.weight"],
                                 e.WS.f32[ENC + "context_encoder.norm.bias"], seed_out=s_pre, training=training,
                                 save_stats=enc_grad)
        if enc_grad:
            c.prenorm = SimpleNamespace(x=xs, m=m, r=r, cfg=cn, seed=s_pre)
        xcur = xn
        executed: List[int] = []
        for j, pre in enumerate(e.block_prefixes):
            is_main = j >= a.prenet_depth
            p_skip = ld_main if is_main else ld_pre
            if p_skip > 0 and not (np.random.random() > p_skip):
                continue
            xcur, t = e._block_forward(e.WS, pre, xcur, B, T, None, training, c.blocks, 16 + 4 * j)
            executed.append(j)
            if is_main:
                layer_results.append(t)
        if not layer_results:
            raise RuntimeError("layerdrop removed every main block of this step")
        top = layer_results[-self.top_k:]
        logits, xmean = ops.layer_mean_head_fwd(top, self.head_w, self.head_b, save_mean=need_grad)
        out: Dict[str, object] = {"encoder_out": logits.view(B, T, self.C), "padding_mask": None,
                                  "layer_results": [t.view(B, T, d) for t in layer_results], "x": xcur.view(B, T, d),
                                  "time_mask": tmask, "channel_mask": cmask, "executed_blocks": executed}
        perm, r_mix = None, 1.0
        if mix and ft.target_mixup:
            perm, r_mix = mix["perm"], mix["r"]
        if target is not None:
            tg = target.to(self.device, torch.float32).contiguous().view(B * T, self.C)
            loss_sum, counters, un, mt = ops.focal_loss_fwd(logits, tg, perm=perm, rows_per_clip=T, r=r_mix,
                                                            threshold=self.metric_threshold,
                                                            want_unreduced=want_unreduced, want_mixed_targets=True)
            out.update(loss_sum=loss_sum, counters=counters, loss_unreduced=un, target=mt.view(B, T, self.C),
                       sample_size=B * T)
            c.tg = tg
        c.logits, c.xmean, c.perm, c.r_mix, c.B, c.T = logits, xmean, perm, r_mix, B, T
        c.n_top, c.executed, c.enc_grad, c.training = len(top), executed, enc_grad, training
        self.ctx = c if need_grad else None
        return out

    # ------------------------------------------------------------------------------------ backward
    def backward(self, grad_scale: Optional[torch.Tensor] = None, dlogits: Optional[torch.Tensor] = None,
                 block_done=None) -> None:
        """Accumulates d(loss_sum * grad_scale) / d(parameters): the head always, the encoder (blocks, prenet norm,
        positional convs, ALiBi scale) once ``freeze_finetune_updates`` has passed. ``dlogits`` (B*T, C) fp32 replaces
        the fused focal-loss gradient when an outer criterion computed the loss from ``encoder_out`` itself."""
        c = self.ctx
        if c is None:
            raise RuntimeError("backward() without a preceding forward(need_grad=True)")
        e, d, B, T = self.core, self.D, c.B, c.T
        if dlogits is None:
            dlogits = ops.focal_loss_bwd(c.logits, c.tg, perm=c.perm, rows_per_clip=T, r=c.r_mix, grad_out=grad_scale)
        else:
            dlogits = dlogits.to(torch.float32).contiguous().view(B * T, self.C)
        g = ops.head_bwd(dlogits, c.xmean, self.head_w, c.n_top, self.head_gw, self.head_gb, want_g=c.enc_grad)
        if not c.enc_grad:
            self.ctx = None
            return
        a = e.a
        main = [j for j in c.executed if j >= a.prenet_depth]
        top_set = set(main[-self.top_k:])
        dx = None
        for s, j in zip(reversed(c.blocks), reversed(c.executed)):
            extra = g if j in top_set else None
            if dx is None and extra is None:
                raise RuntimeError("internal: block without gradient")  # cannot happen: the last executed main block is in top_set
            dx = e._block_backward(e.WS, s, dx, B, T, None, c.training, extra_dt=extra)
            if block_done is not None:
                block_done(j)
        c.blocks = None
        s = c.prenorm
        dxs, _ = ops.rowln_bwd(s.cfg, dx, s.x, None, e.WS.f32[ENC + "context_encoder.norm.weight"],
                               e.WS.f32[ENC + "context_encoder.norm.bias"], None, None, s.m, s.r, seed_out=s.seed,
                               training=c.training, dgamma=e.G(ENC + "context_encoder.norm.weight"),
                               dbeta=e.G(ENC + "context_encoder.norm.bias"))
        # xs = x + x_pos(x): d x = d xs + dgrad of the positional stack; x = channel_mask(time_mask(local_features))
        dpos = dxs.view(B, T, d)
        gq = a.conv_pos_groups
        cfg_l = ops.RowLnCfg(d, 1e-5, act=1)
        for li in reversed(range(len(e.pos_names))):
            sp = c.pos[li]
            n = e.pos_names[li]
            dy, _ = ops.rowln_bwd(cfg_l, dpos, sp.y, None, None, None, None, None, sp.m, sp.r)
            ops.colsum(dy.view(B * T, d), e.G(n[:-6] + "bias"))
            e.conv_wgrad(dy, sp.x, e.gpacked[n + "|F"], taps=e.kp, pad=e.kp // 2, groups=gq)
            dpos = e.conv(dy, e.WS, n, taps=e.kp, pad=e.kp - 1 - e.kp // 2, groups=gq, dgrad=True)
            c.pos[li] = None
        dx_in = e._add(dpos.view(B * T, d), dxs)
        if getattr(c, "cm_d", None) is not None:  # masked channels were overwritten with 0: no gradient
            ops.channel_mask_(dx_in, c.cm_d, T)
        if getattr(c, "keep_idx", None) is not None:  # masked frames were replaced by noise tokens: no gradient
            dx_in = ops.row_gather(dx_in, c.keep_idx, B * T)
        e._proj_backward(c, dx_in)  # project_features trains; the gradient stops at the frozen conv extractor
        e._unpack_grads()
        self.ctx = None
