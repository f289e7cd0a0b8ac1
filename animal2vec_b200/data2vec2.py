"""Drop-in ``Data2VecMultiModel`` for the pretraining path (/root/reference/nn/data2vec2.py:168-1150).

Same registry name (``data2vec_multi``), constructor / ``build_model`` signature, ``forward`` keyword
surface, ``set_num_updates``, result-dict keys and state-dict keys (checkpoint ABI),
but every device operation runs in the hand-written sm_100a kernels scheduled by
:class:`animal2vec_b200.engine.PretrainEngine`. There is no PyTorch-eager or CPU implementation behind
this module: constructing it without the CUDA library, or calling it with CPU tensors, raises.

The sub-module tree (``modality_encoders.AUDIO.local_encoder.conv_layers.0.0`` ...) exists to carry the
parameters under the reference's names; parameters are views into the engine's flat fp32 buffer and
``p.grad`` are views into its flat gradient buffer, so an outer trainer / optimizer (fairseq's included)
sees ordinary ``nn.Parameter`` objects.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import registry
from .config import Data2VecMultiConfig, Modality, from_dict, resolve
from .engine import PretrainEngine, annealed_decay


class _Holder(nn.Module):
    """Parameter container node (the arithmetic lives in the fused engine, not in per-module forwards)."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("this sub-module only names parameters; the computation is fused inside "
                           "Data2VecMultiModel.forward (animal2vec_b200.engine)")


class _StepFunction(torch.autograd.Function):
    """Single autograd node of the whole pretraining forward: backward runs the engine's kernel schedule and
    accumulates into ``p.grad`` (views of the flat gradient buffer) as a side effect."""

    @staticmethod
    def forward(ctx, anchor, model, loss_sum):
        ctx.model = model
        return loss_sum.clone()

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.model
        g = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        m._attach_grads()
        m.engine.backward(g, training=m.training)
        return None, None, None


@registry.register_model("data2vec_multi", dataclass=Data2VecMultiConfig)
class Data2VecMultiModel(nn.Module):
    def __init__(self, cfg: Data2VecMultiConfig, modalities=None, skip_ema=False, task=None, *,
                 precision: str = "bf16", device="cuda", init: Optional[Dict[str, torch.Tensor]] = None,
                 init_seed: int = 0):
        super().__init__()
        if isinstance(cfg, dict):
            cfg = from_dict(Data2VecMultiConfig, cfg)
        self.cfg = resolve(cfg)
        self.modalities = modalities if modalities is not None else [Modality.AUDIO]
        self.task = task
        if skip_ema or cfg.skip_ema:
            raise NotImplementedError("skip_ema: the teacher is part of the fused pretraining step")
        self.engine = PretrainEngine(self.cfg, device, precision=precision, init=init, init_seed=init_seed,
                                     rng_seed=int(self.cfg.seed))
        self.num_updates = 0
        self._params: Dict[str, nn.Parameter] = {}
        for name in self.engine.S.shapes:  # reference named_parameters() order
            p = nn.Parameter(self.engine.S.view(name), requires_grad=True)
            # nn/data2vec2.py:318-322
            if len(p.shape) == 1 or name.endswith(".bias") or "alibi_scale" in name or "p_swish" in name:
                p.optim_overrides = {"optimizer": {"weight_decay_scale": 0}}
            if self.cfg.decoder_group and "decoder" in name:
                p.param_group = "decoder"
            self._register(name, p)
            self._params[name] = p
        self._anchor = torch.zeros(1, device=self.engine.device, requires_grad=True)
        self._attach_grads()
        self._pending_guard = None

    # ---------------------------------------------------------------------------------- plumbing
    def _register(self, dotted: str, p: nn.Parameter) -> None:
        node: nn.Module = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            child = node._modules.get(part)
            if child is None:
                child = _Holder()
                node.add_module(part, child)
            node = child
        node.register_parameter(parts[-1], p)

    def _attach_grads(self) -> None:
        for name, p in self._params.items():
            g = self.engine.S.gview(name)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    @classmethod
    def build_model(cls, cfg: Data2VecMultiConfig, task=None):
        """nn/data2vec2.py:500-514."""
        if task is None or not hasattr(task, "supported_modalities"):
            modalities = [cfg.supported_modality] if cfg.supported_modality is not None else [Modality.AUDIO]
        else:
            modalities = task.supported_modalities
        return cls(cfg, modalities, task=task)

    # ---------------------------------------------------------------------------------- EMA / updates
    def set_num_updates(self, num_updates: int) -> None:
        """nn/data2vec2.py:386-410. Called by the trainer after each optimizer step: the student's fp32
        masters were modified in place by the optimizer, so the GEMM operand copies are rebuilt lazily."""
        self.engine.mark_student_updated()
        if (self.num_updates == 0 and num_updates > 1) or self.num_updates >= num_updates:
            pass  # checkpoint restore / repeated call: no EMA step (reference :389-393)
        elif self.training:
            self.engine.ema_step(num_updates)
        self.num_updates = num_updates

    # ---------------------------------------------------------------------------------- checkpoint ABI
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        state = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        state[prefix + "_ema"] = {k: self.engine.E.view(k).detach().clone() for k in self.engine.E.names}
        return state

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        state_dict = dict(state_dict)
        ema = state_dict.pop("_ema", None)
        out = super().load_state_dict(state_dict, strict=strict)
        self.engine.mark_student_updated()
        if ema is not None:
            self.engine.load_teacher(ema)
        else:
            self.engine.reset_teacher()
        return out

    # ---------------------------------------------------------------------------------- forward
    def forward(self, source, target=None, id=None, mode=None, padding_mask=None, mask=True, features_only=False,
                force_remove_masked=False, remove_extra_tokens=True, precomputed_mask=None, reduce=True, **kwargs):
        """Pretraining branch of nn/data2vec2.py:516-991. Returns the reference's result dict; the
        ``AUDIO_regression`` entry is the already-summed loss as a 1-element fp32 tensor (the criterion's
        ``.float().sum()`` and ``backward()`` work unchanged; the unreduced (N_masked, D) tensor is never
        materialised)."""
        if features_only or not mask:
            raise NotImplementedError("features_only / mask=False (finetune + inference path) is a 'next' row of "
                                      "SURVEY.md section 8(f); the pretraining path is implemented")
        if padding_mask is not None:
            raise NotImplementedError("padding_mask: the reference's own convert_padding_mask is broken "
                                      "(nn/modalities/audio.py:168); the shipped task disables padding")
        if mode is not None and (mode.name if isinstance(mode, Modality) else str(mode)) != "AUDIO":
            raise NotImplementedError("only the AUDIO modality exists on this path")
        cfg, e = self.cfg, self.engine
        self._check_guards()
        need_grad = torch.is_grad_enabled() and self.training
        mask_np = None
        if precomputed_mask is not None:
            mask_np = precomputed_mask.detach().bool().cpu().numpy()
        if need_grad:
            self._attach_grads()
        res = e.forward(source, id, self.num_updates, mask=mask_np, training=self.training, need_grad=need_grad)
        loss_sum = res["loss_sum"].to(torch.float32)
        if need_grad:
            loss_t = _StepFunction.apply(self._anchor, self, loss_sum)
        else:
            loss_t = loss_sum
        n = res["sample_size"]
        pred_var, target_var = PretrainEngine.variances(res["colstats"], n)
        result = {
            "losses": {"AUDIO_regression": loss_t},
            "sample_size": torch.tensor(n, dtype=torch.long, device=e.device),
            "masked_pct": res["masked_pct"],
            "pred_var": pred_var.float(),
            "target_var": target_var.float(),
            "ema_decay": annealed_decay(cfg, self.num_updates) * 1000,
        }
        if self.num_updates > 5000:
            self._pending_guard = (pred_var, target_var)
        return result

    def _check_guards(self) -> None:
        """Representation-collapse guards of nn/data2vec2.py:972-988, evaluated one step late so that the
        comparison never stalls the GPU queue."""
        g, self._pending_guard = self._pending_guard, None
        if g is None:
            return
        pv, tv = float(g[0]), float(g[1])
        if tv < self.cfg.min_target_var:
            raise Exception(f"target var is {tv} < {self.cfg.min_target_var}, exiting ({'AUDIO'})")
        if pv < self.cfg.min_pred_var:
            raise Exception(f"pred var is {pv} < {self.cfg.min_pred_var}, exiting ({'AUDIO'})")

    def extract_features(self, source, mode=None, padding_mask=None, mask=False, remove_extra_tokens=True):
        return self.forward(source, mode=mode, padding_mask=padding_mask, mask=mask, features_only=True,
                            remove_extra_tokens=remove_extra_tokens)

    def remove_pretraining_modules(self, modality=None, keep_decoder=False):
        raise NotImplementedError("finetune hand-over is a 'next' row (SURVEY.md section 8f-1)")
