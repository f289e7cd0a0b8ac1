"""Name registries mirroring fairseq's ``@register_model / @register_task / @register_criterion``.

The reference registers its classes with fairseq (nn/data2vec2.py:168, nn/wav2vec2.py:57,
nn/audio_tasks.py:92, nn/criterions.py:137,388). fairseq is not a dependency of this package. When it IS
importable, the classes below derive from fairseq's own base classes (:func:`fairseq_bases`) and are
registered with fairseq's registrars under the same names, so a fairseq/hydra config that names
``data2vec_multi`` / ``expanded_model`` / ``audio_ccas`` resolves to the B200 implementations. A registrar
error (duplicate name, wrong base class) is raised, never swallowed. The fairseq side of this hook is
exercised against a stub package in tests/test_host_cpu.py (fairseq itself cannot be installed here).
"""
from __future__ import annotations

import importlib
from typing import Callable, Dict, Tuple

MODELS: Dict[str, type] = {}
CRITERIA: Dict[str, type] = {}
TASKS: Dict[str, type] = {}
DATACLASSES: Dict[str, type] = {}
FAIRSEQ_REGISTERED: Dict[str, str] = {}  # "<kind>:<name>" -> class name, filled when fairseq took the class

_KIND_MODULE = {"model": ("fairseq.models", "register_model"),
                "criterion": ("fairseq.criterions", "register_criterion"),
                "task": ("fairseq.tasks", "register_task")}


def fairseq_available() -> bool:
    try:
        importlib.import_module("fairseq")
        return True
    except ImportError:
        return False


def fairseq_bases() -> Tuple[type, type, type]:
    """(model base, criterion base, task base): fairseq's when importable, else torch.nn.Module / object."""
    import torch.nn as nn

    if not fairseq_available():
        return nn.Module, nn.Module, object
    from fairseq.criterions import FairseqCriterion
    from fairseq.models import BaseFairseqModel
    from fairseq.tasks import FairseqTask

    return BaseFairseqModel, FairseqCriterion, FairseqTask


def _register(table: Dict[str, type], kind: str, name: str, dataclass) -> Callable[[type], type]:
    def deco(cls: type) -> type:
        if name in table and table[name] is not cls:
            raise ValueError(f"{kind} {name!r} already registered")
        table[name] = cls
        if dataclass is not None:
            DATACLASSES[name] = dataclass
        cls._registry_name = name
        if fairseq_available():
            mod_name, fn_name = _KIND_MODULE[kind]
            registrar = getattr(importlib.import_module(mod_name), fn_name)
            registrar(name, dataclass=dataclass)(cls)  # raises on a duplicate name / wrong base class
            FAIRSEQ_REGISTERED[f"{kind}:{name}"] = cls.__name__
        return cls

    return deco


def register_model(name: str, dataclass=None):
    return _register(MODELS, "model", name, dataclass)


def register_criterion(name: str, dataclass=None):
    return _register(CRITERIA, "criterion", name, dataclass)


def register_task(name: str, dataclass=None):
    return _register(TASKS, "task", name, dataclass)
