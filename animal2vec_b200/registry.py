"""Name registries mirroring fairseq's ``@register_model / @register_task / @register_criterion``.

The reference registers its classes with fairseq (nn/data2vec2.py:168, nn/audio_tasks.py:92,
nn/criterions.py:388). fairseq is not a dependency of this package; when it IS importable the same
classes are additionally registered with it under the same names, so a fairseq/hydra config that names
``data2vec_multi`` / ``expanded_model`` resolves to the B200 implementations.
"""
from __future__ import annotations

from typing import Callable, Dict

MODELS: Dict[str, type] = {}
CRITERIA: Dict[str, type] = {}
TASKS: Dict[str, type] = {}
DATACLASSES: Dict[str, type] = {}


def _also_fairseq(kind: str, name: str, dataclass):
    try:  # pragma: no cover - fairseq is absent in the build image
        import fairseq  # noqa: F401
        from fairseq import criterions, models, tasks

        return {"model": models.register_model, "criterion": criterions.register_criterion,
                "task": tasks.register_task}[kind](name, dataclass=dataclass)
    except Exception:
        return None


def _register(table: Dict[str, type], kind: str, name: str, dataclass) -> Callable[[type], type]:
    def deco(cls: type) -> type:
        if name in table and table[name] is not cls:
            raise ValueError(f"{kind} {name!r} already registered")
        table[name] = cls
        if dataclass is not None:
            DATACLASSES[name] = dataclass
        cls._registry_name = name
        return cls

    return deco


def register_model(name: str, dataclass=None):
    return _register(MODELS, "model", name, dataclass)


def register_criterion(name: str, dataclass=None):
    return _register(CRITERIA, "criterion", name, dataclass)


def register_task(name: str, dataclass=None):
    return _register(TASKS, "task", name, dataclass)
