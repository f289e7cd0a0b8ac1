"""Drop-in finetune model ``wav2vec_ccas_finetune`` (/root/reference/nn/wav2vec2.py:57-482): same registry name,
config dataclass, ``build_model`` / ``forward`` / ``get_logits`` / ``get_targets`` / ``set_num_updates`` surface and
state-dict keys (``w2v_encoder.w2v_model.*`` + ``w2v_encoder.proj.*``); the arithmetic is the kernel schedule of
:class:`animal2vec_b200.finetune.FinetuneEngine`. The whole forward is ONE autograd node: ``encoder_out`` (B x T x C
logits) carries the graph, its backward runs the engine's backward with the incoming ``dlogits``.
"""
from __future__ import annotations

import ast
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import registry
from .config import Data2VecMultiConfig, Wav2Vec2CcasFinetuneConfig, from_dict
from .data2vec2 import _Holder
from .finetune import FinetuneEngine

_ModelBase = registry.fairseq_bases()[0]


class _FinetuneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, encoder, logits):
        ctx.encoder = encoder
        return logits.clone()

    @staticmethod
    def backward(ctx, grad_out):
        enc = ctx.encoder
        enc._attach_grads()
        enc.engine.backward(dlogits=grad_out.detach())
        return None, None, None


class Wav2VecEncoderModOut(nn.Module):
    """nn/wav2vec2.py:93-482. ``w2v_model`` names the pretrained parameters (views of the engine's flat buffer),
    ``proj`` the classification head."""

    def __init__(self, cfg: Wav2Vec2CcasFinetuneConfig, output_size: int, *, model_cfg: Optional[Data2VecMultiConfig] = None,
                 state: Optional[Dict] = None, precision: str = "bf16", device="cuda", metric_threshold: float = 0.25):
        super().__init__()
        if model_cfg is None:
            w2v_args = cfg.w2v_args
            model_cfg = w2v_args.get("model") if isinstance(w2v_args, dict) else getattr(w2v_args, "model", None)
            if model_cfg is None:
                raise ValueError("w2v_args.model (the pretraining model config) or model_cfg= is required; loading it "
                                 "from cfg.w2v_path needs fairseq's checkpoint_utils (nn/wav2vec2.py:132-141)")
        if isinstance(model_cfg, dict):
            model_cfg = from_dict(Data2VecMultiConfig, model_cfg)
        if not cfg.normalize:
            raise AssertionError("data2vec_multi finetuning asserts cfg.normalize (nn/wav2vec2.py:178)")
        init = None
        if state is not None and not cfg.no_pretrained_weights:
            init = self._pretrained_weights(state, cfg)
        self.cfg = cfg
        self.engine = FinetuneEngine(model_cfg, cfg, output_size, device, precision=precision, init=init,
                                     metric_threshold=metric_threshold)
        self.apply_mask = cfg.apply_mask
        self.freeze_finetune_updates = cfg.freeze_finetune_updates
        self.num_updates = 0
        self.w2v_model = _Holder()
        self._params: Dict[str, nn.Parameter] = {}
        core = self.engine.core
        for name in core.S.shapes:
            if ".decoder." in name:  # remove_pretraining_modules dropped it (nn/data2vec2.py:1125-1142)
                continue
            p = nn.Parameter(core.S.view(name), requires_grad=True)
            if len(p.shape) == 1 or name.endswith(".bias") or "alibi_scale" in name or "p_swish" in name:
                p.optim_overrides = {"optimizer": {"weight_decay_scale": 0}}
            self._register(self.w2v_model, name, p)
            self._params["w2v_model." + name] = p
        self.proj = _Holder()
        self.proj.register_parameter("weight", nn.Parameter(self.engine.head_w))
        self.proj.register_parameter("bias", nn.Parameter(self.engine.head_b))
        self._apply_layer_decay(getattr(cfg, "layer_decay", 1))
        self._anchor = torch.zeros(1, device=self.engine.device, requires_grad=True)
        self._attach_grads()

    @staticmethod
    def _pretrained_weights(state: Dict, cfg: Wav2Vec2CcasFinetuneConfig) -> Dict[str, torch.Tensor]:
        """nn/wav2vec2.py:190-197 (load_ema) + load_model_weights :311-360: the EMA teacher's weights replace the
        student's when ``load_ema``; ``_ema`` / decoder entries are dropped."""
        sd = dict(state["model"] if "model" in state else state)
        ema = sd.pop("_ema", None)
        if cfg.load_ema:
            assert ema is not None, "_ema"
            for k, v in ema.items():
                assert k in sd, k
                sd[k] = v
        return {k: v for k, v in sd.items() if torch.is_tensor(v)}

    @staticmethod
    def _register(root: nn.Module, dotted: str, p: nn.Parameter) -> None:
        node = root
        parts = dotted.split(".")
        for part in parts[:-1]:
            child = node._modules.get(part)
            if child is None:
                child = _Holder()
                node.add_module(part, child)
            node = child
        node.register_parameter(parts[-1], p)

    def _apply_layer_decay(self, layer_decay: float) -> None:
        """nn/wav2vec2.py:213-233: lr_scale = layer_decay ** (num_layers - i) on the block parameters."""
        if layer_decay >= 1:
            return
        prefixes = self.engine.core.block_prefixes
        num_layers = len(prefixes) + 1
        scales = [layer_decay ** (num_layers - i) for i in range(num_layers + 1)]
        for i, pre in enumerate(prefixes):
            if scales[i + 1] == 1.0:
                continue
            for name, p in self._params.items():
                if name.startswith("w2v_model." + pre):
                    ov = dict(getattr(p, "optim_overrides", {}))
                    ov.setdefault("optimizer", {})
                    ov["optimizer"] = dict(ov["optimizer"], lr_scale=scales[i + 1])
                    p.optim_overrides = ov

    def _attach_grads(self) -> None:
        core = self.engine.core
        missing = [n for n, p in self._params.items() if p.grad is None]
        if missing:
            if len(missing) == len(self._params):
                core.zero_grad()
            else:
                for n in missing:
                    core.S.gview(n[len("w2v_model."):]).zero_()
        for name, p in self._params.items():
            g = core.S.gview(name[len("w2v_model."):])
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
        for p, g in ((self.proj.weight, self.engine.head_gw), (self.proj.bias, self.engine.head_gb)):
            if p.grad is None:
                g.zero_()
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.engine.zero_grad()
        self._attach_grads()

    def _apply(self, fn, recurse=True):
        probe = torch.empty(0, dtype=torch.float32, device=self.engine.device)
        out = fn(probe)
        if out.dtype != probe.dtype or out.device != probe.device:
            raise RuntimeError("finetune parameters are views of the CUDA engine's fp32 master buffers: casting or "
                               "moving the module is not supported (precision is chosen at construction)")
        return self

    def set_num_updates(self, num_updates: int) -> None:
        self.num_updates = num_updates
        self.engine.num_updates = num_updates
        self.engine.core.mark_student_updated()

    def forward(self, source, padding_mask=None, target=None, **kwargs):
        if padding_mask is not None:
            raise NotImplementedError("padding_mask (the shipped task disables padding)")
        need_grad = torch.is_grad_enabled() and self.training
        if need_grad:
            self._attach_grads()
        res = self.engine.forward(source, target, training=self.training, need_grad=need_grad)
        logits = res["encoder_out"]
        if need_grad:
            logits = _FinetuneFunction.apply(self._anchor, self, logits)
        return {"encoder_out": logits, "padding_mask": None, "layer_results": res["layer_results"],
                "target": res.get("target", target), "_fused": res}


@registry.register_model("wav2vec_ccas_finetune", dataclass=Wav2Vec2CcasFinetuneConfig)
class Wav2VecCcasFinetune(_ModelBase):
    def __init__(self, cfg: Wav2Vec2CcasFinetuneConfig, w2v_encoder: Wav2VecEncoderModOut):
        super().__init__()
        self.cfg = cfg
        self.w2v_encoder = w2v_encoder

    @classmethod
    def build_model(cls, cfg: Wav2Vec2CcasFinetuneConfig, task=None, **kw):
        """nn/wav2vec2.py:59-66."""
        if isinstance(cfg, dict):
            cfg = from_dict(Wav2Vec2CcasFinetuneConfig, cfg)
        labels = ast.literal_eval(cfg.unique_labels)
        return cls(cfg, Wav2VecEncoderModOut(cfg, len(labels), **kw))

    def forward(self, **kwargs):
        return self.w2v_encoder(**kwargs)

    def set_num_updates(self, num_updates: int) -> None:
        self.w2v_encoder.set_num_updates(num_updates)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.w2v_encoder.zero_grad()

    def _apply(self, fn, recurse=True):
        self.w2v_encoder._apply(fn)
        return self

    def get_logits(self, net_output, reshape=True):
        y = net_output["encoder_out"]
        return y.reshape(-1, y.size(-1)) if reshape else y

    @staticmethod
    def prepare_shapes(out, transpose=True):
        if transpose:
            out = out.transpose(0, 2)
        return out.reshape(-1, out.size(-1))

    def get_targets(self, sample, net_output, expand_steps=True, reshape=True):
        y = net_output["target"] if net_output.get("target") is not None else sample["target"]
        if reshape:
            y = self.prepare_shapes(y, transpose=False) if self.cfg.use_focal_loss else y.reshape(-1)
        return y

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        logits = net_output["encoder_out"].float()
        return torch.log_softmax(logits, -1) if log_probs else torch.softmax(logits, -1)

    def prepare_for_inference_(self, cfg=None):
        self.eval()

    def max_positions(self):
        return None
