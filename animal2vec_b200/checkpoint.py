"""Reference checkpoint compatibility.

The reference's trainer (fairseq) writes ``{"model": state_dict incl. "_ema", "cfg": {"model": ..., "task": ...}, ...}``
(README.md:42-44; nn/data2vec2.py:412-429; nn/wav2vec2.py:132-141 reads it back). This module turns such a dict (already
``torch.load``-ed; omegaconf nodes, plain dicts and argparse-style namespaces are accepted) into this package's config
dataclass and tensors: field names are the reference's, ``alibi_scale`` tensors written before the layer axis existed are
upgraded (nn/modalities/base.py:152-157), ``_ema`` feeds the fp32 teacher shadow.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch

from .config import Data2VecMultiConfig, Modality, from_dict, resolve
from .params import ENC, is_teacher_key, student_param_shapes


def _plain(node: Any) -> Any:
    """omegaconf DictConfig / Namespace / dict -> plain nested dict."""
    if node is None or isinstance(node, (int, float, str, bool)):
        return node
    if hasattr(node, "items"):
        return {str(k): _plain(v) for k, v in node.items()}
    if hasattr(node, "__dict__") and not torch.is_tensor(node):
        return {k: _plain(v) for k, v in vars(node).items()}
    if isinstance(node, (list, tuple)):
        return [_plain(v) for v in node]
    return node


def model_config_from_checkpoint(state: Dict[str, Any]) -> Data2VecMultiConfig:
    """``state["cfg"]["model"]`` (+ the task fields the model interpolates) -> Data2VecMultiConfig."""
    cfg = _plain(state.get("cfg") or {})
    model = dict(cfg.get("model") or {})
    if not model:
        raise KeyError("checkpoint has no cfg.model node")
    name = model.pop("_name", "data2vec_multi")
    if "data2vec_multi" not in str(name):
        raise ValueError(f"checkpoint model is {name!r}, expected data2vec_multi")
    task = cfg.get("task") or {}
    known = {f for f in Data2VecMultiConfig.__dataclass_fields__}
    model = {k: v for k, v in model.items() if k in known}
    if isinstance(model.get("supported_modality"), str):
        model["supported_modality"] = Modality[model["supported_modality"]]
    audio = (model.get("modalities") or {}).get("audio")
    if isinstance(audio, dict) and isinstance(audio.get("type"), str):
        audio["type"] = Modality[audio["type"]]
    out = from_dict(Data2VecMultiConfig, model)
    for k in ("sample_rate", "conv_feature_layers"):  # II("task.*") interpolations
        if task.get(k) is not None and getattr(out, k, None) in (None, "???"):
            setattr(out, k, task[k])
    return resolve(out)


def split_model_state(state: Dict[str, Any]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(student tensors, teacher shadow tensors) of a checkpoint dict or bare state dict, alibi_scale upgraded."""
    sd = dict(state["model"] if ("model" in state and isinstance(state["model"], dict)) else state)
    ema = dict(sd.pop("_ema", None) or {})
    for d in (sd, ema):
        k = ENC + "alibi_scale"
        if k in d and d[k].dim() == 4:
            d[k] = d[k].unsqueeze(0)
    student = {k: v for k, v in sd.items() if torch.is_tensor(v)}
    return student, ema


def check_inventory(cfg: Data2VecMultiConfig, student: Dict[str, torch.Tensor], ema: Dict[str, torch.Tensor]) -> None:
    """The checkpoint ABI: every key / shape the engine expects (and nothing else under the model prefix)."""
    shapes = student_param_shapes(cfg)
    missing = sorted(set(shapes) - set(student))
    extra = sorted(k for k in set(student) - set(shapes) if not k.startswith("_"))
    bad = sorted(k for k in shapes if k in student and tuple(student[k].shape) != tuple(shapes[k]))
    if missing or extra or bad:
        raise KeyError(f"checkpoint does not match the model: missing {missing[:5]}, unexpected {extra[:5]}, shape {bad[:5]}")
    if ema:
        want = {k for k in shapes if is_teacher_key(k)}
        if set(ema) != want:
            raise KeyError(f"_ema keys differ: missing {sorted(want - set(ema))[:5]}, unexpected {sorted(set(ema) - want)[:5]}")


def load_pretrained(state: Dict[str, Any], **model_kw):
    """Data2VecMultiModel from a reference checkpoint dict (needs the CUDA library and a B200)."""
    from .data2vec2 import Data2VecMultiModel

    cfg = model_config_from_checkpoint(state)
    student, ema = split_model_state(state)
    check_inventory(cfg, student, ema)
    model = Data2VecMultiModel(cfg, init=student, **model_kw)
    if ema:
        model.engine.load_teacher(ema)
    return model
