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

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import registry
from .config import Data2VecMultiConfig, Modality, from_dict, resolve
from .engine import ENC, PretrainEngine, annealed_decay

# fairseq's BaseFairseqModel when fairseq is importable (then the class is also registered with fairseq under the
# reference's name, see registry._register), torch.nn.Module otherwise
_ModelBase = registry.fairseq_bases()[0]


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
class Data2VecMultiModel(_ModelBase):
    def __init__(self, cfg: Data2VecMultiConfig, modalities=None, skip_ema=False, task=None, *,
                 precision: Optional[str] = None, device="cuda", init: Optional[Dict[str, torch.Tensor]] = None,
                 init_seed: int = 0):
        super().__init__()
        if precision is None:  # build_model(cfg, task) has no precision argument: environment switch
            precision = os.environ.get("A2V_PRECISION", "bf16")
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
            if name.endswith("alibi_scale") and not self.cfg.modalities.audio.learned_alibi_scale:
                p.requires_grad_(False)  # nn/modalities/base.py:134
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
        """``p.grad`` are views of the engine's flat gradient buffer. An outer optimizer's ``zero_grad`` (fairseq's
        FairseqOptimizer.zero_grad, torch's ``set_to_none=True`` default) only drops the views: whenever a view is
        found missing its slice of the flat buffer is cleared before it is attached again, so gradients never pile
        up across optimizer steps."""
        missing = [n for n, p in self._params.items() if p.grad is None]
        if missing:
            if len(missing) == len(self._params):
                self.engine.zero_grad()
            else:
                for n in missing:
                    self.engine.S.gview(n).zero_()
        for name, p in self._params.items():
            g = self.engine.S.gview(name)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                if p.grad is not None:  # a foreign gradient tensor was assigned: keep its values
                    g.copy_(p.grad)
                p.grad = g

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Clears the flat gradient buffer in one pass (the views stay attached regardless of ``set_to_none``)."""
        self.engine.zero_grad()
        self._attach_grads()

    def _apply(self, fn, recurse=True):
        """Parameters alias the engine's flat fp32 master buffer: ``.half()`` / ``.bfloat16()`` / ``.to(dtype)`` /
        ``.cpu()`` would silently replace them with detached copies the kernels never read (fairseq calls
        ``model.half()`` under ``common.fp16``). The arithmetic precision is chosen through ``precision=`` (bf16
        kernels with fp32 masters); dtype or device moves are refused, same-device no-ops are accepted."""
        probe = torch.empty(0, dtype=torch.float32, device=self.engine.device)
        out = fn(probe)
        if out.dtype != probe.dtype or out.device != probe.device:
            raise RuntimeError(
                "Data2VecMultiModel parameters are views of the CUDA engine's fp32 master buffer: casting or moving "
                f"the module ({probe.dtype}/{probe.device} -> {out.dtype}/{out.device}) is not supported; choose "
                "the compute precision with precision='bf16'|'fp32' at construction (no model.half())")
        return self

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
        state = nn.Module.state_dict(self, *args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        if self.engine.has_teacher:
            state[prefix + "_ema"] = {k: self.engine.E.view(k).detach().clone() for k in self.engine.E.names}
        return state

    def upgrade_state_dict_named(self, state_dict, name=""):
        """nn/modalities/base.py:152-157: checkpoints written before ``alibi_scale`` gained its leading layer axis
        hold a 4-D tensor. Applied to the student keys and to the ``_ema`` dict."""
        pre = (name + ".") if name else ""
        for sd in (state_dict, state_dict.get(pre + "_ema") if isinstance(state_dict.get(pre + "_ema"), dict) else None):
            if sd is None:
                continue
            for k in (pre + ENC + "alibi_scale", ENC + "alibi_scale"):
                if k in sd and torch.is_tensor(sd[k]) and sd[k].dim() == 4:
                    sd[k] = sd[k].unsqueeze(0)
        return state_dict

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False, model_cfg=None, args=None):
        """nn/data2vec2.py:420-429 (``_ema`` consumed into the fp32 teacher shadow) + fairseq's
        BaseFairseqModel.load_state_dict (upgrade hook first). Accepts the reference's checkpoint layouts: a plain
        state dict, or the trainer's ``{"model": ..., "cfg": ...}`` file dict."""
        if "model" in state_dict and isinstance(state_dict["model"], dict) and "_ema" not in state_dict:
            state_dict = state_dict["model"]
        state_dict = dict(state_dict)
        self.upgrade_state_dict_named(state_dict, "")
        ema = state_dict.pop("_ema", None)
        if self.engine.has_teacher and ema is None and strict:
            raise KeyError("_ema")  # the reference asserts the key (nn/data2vec2.py:422-423)
        if not self.engine.has_decoder:  # wav2vec2.py:332-335: decoder weights of a pretraining checkpoint are dropped
            state_dict = {k: v for k, v in state_dict.items() if not k.startswith(ENC + "decoder.")}
        out = nn.Module.load_state_dict(self, state_dict, strict=strict)
        self.engine.mark_student_updated()
        if self.engine.has_teacher:
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
        if padding_mask is not None:
            raise NotImplementedError("padding_mask: the reference's own convert_padding_mask is broken "
                                      "(nn/modalities/audio.py:168); the shipped task disables padding")
        if mode is not None and (mode.name if isinstance(mode, Modality) else str(mode)) != "AUDIO":
            raise NotImplementedError("only the AUDIO modality exists on this path")
        cfg, e = self.cfg, self.engine
        if features_only:
            # nn/data2vec2.py:632-728 with features_only=True: clone_batch 1, masked rows stay in place
            # (remove_masked = force_remove_masked, unsupported), no decoder, no teacher
            if force_remove_masked:
                raise NotImplementedError("force_remove_masked on the features_only path")
            mask_np = None if precomputed_mask is None else precomputed_mask.detach().bool().cpu().numpy()
            res = e.extract_features(source, ids=id, num_updates=self.num_updates, mask=bool(mask),
                                     precomputed_mask=mask_np, training=self.training, need_grad=False)
            return {"x": res["x"], "linear_eval_projection": None, "padding_mask": None,
                    "layer_results": res["layer_results"], "mask": res["mask"]}
        if not mask:
            raise NotImplementedError("mask=False without features_only: the pretraining loss needs masked rows")
        if not (e.has_teacher and e.has_decoder):
            raise RuntimeError("remove_pretraining_modules() was called: only features_only=True forwards remain")
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
        stats = res["colstats"]
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1:
            # compute_var (nn/data2vec2.py:1095-1105) all-reduces count, sum and sum of squares: one packed collective
            packed = torch.cat([stats.reshape(-1), torch.tensor([float(n)], device=stats.device, dtype=stats.dtype)])
            torch.distributed.all_reduce(packed)
            stats, n_var = packed[:-1].view_as(stats), packed[-1]
        else:
            n_var = n
        pred_var, target_var = PretrainEngine.variances(stats, n_var)
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
        """nn/data2vec2.py:1125-1142: drop the EMA teacher, set clone_batch 1 and (unless ``keep_decoder``) the
        decoder. Afterwards only ``features_only=True`` forwards (extract_features) are possible and the state
        dict carries neither ``_ema`` nor decoder keys."""
        if modality is not None and str(modality).lower() != "audio":
            raise NotImplementedError("only the AUDIO modality exists on this path")
        self.engine.drop_teacher()
        self.cfg.clone_batch = 1
        if not keep_decoder and self.engine.has_decoder:
            self.engine.drop_decoder()
            enc = self.modality_encoders.AUDIO
            if "decoder" in enc._modules:
                del enc._modules["decoder"]
            for n in [n for n in self._params if n.startswith(ENC + "decoder.")]:
                del self._params[n]

    # ---------------------------------------------------------------------------------- fairseq model protocol
    def prepare_for_inference_(self, cfg=None):
        self.eval()

    def max_positions(self):
        return None

    def get_targets(self, sample, net_output):
        return sample.get("target") if isinstance(sample, dict) else None

    def get_logits(self, net_output, reshape=True):
        y = net_output["linear_eval_projection"]
        if y is None:
            raise NotImplementedError("with_labels (linear-eval projection during pretraining) is not on this path")
        return y.reshape(-1, y.size(-1)) if reshape else y
