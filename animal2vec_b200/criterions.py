"""``ExpandedModelCriterion`` (registry name ``expanded_model``), the pretraining criterion of
/root/reference/nn/criterions.py:388-502. Its forward is fairseq's ``ModelCriterion.forward`` (third
party; behaviour restated in SURVEY.md Appendix B3): sum the model-supplied losses, take the model's
``sample_size``, copy the requested log keys."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import registry


@dataclass
class ExpandedModelCriterionConfig:
    """nn/criterions.py:83-134 (+ fairseq ModelCriterionConfig: loss_weights, log_keys)."""

    loss_weights: Dict[str, float] = field(default_factory=dict)
    log_keys: List[str] = field(default_factory=list)
    unique_labels: Optional[str] = None
    verbose_tensorboard_logging: bool = False
    segmentation_metrics: bool = False
    use_focal_loss: bool = True
    metric_threshold: float = 0.25
    iou_threshold: float = 0.0
    sigma_s: float = 0.1
    maxfilt_s: float = 0.1
    max_duration_s: float = 0.5
    lowP: float = 0.125
    method: str = "avg"
    can_sum: bool = True


_CriterionBase = registry.fairseq_bases()[1]  # fairseq's FairseqCriterion when importable, else torch.nn.Module


def _init_base(obj, task) -> None:
    if _CriterionBase is torch.nn.Module:
        torch.nn.Module.__init__(obj)
    else:
        _CriterionBase.__init__(obj, task)
    obj.task = task


@registry.register_criterion("expanded_model", dataclass=ExpandedModelCriterionConfig)
class ExpandedModelCriterion(_CriterionBase):
    def __init__(self, task, loss_weights=None, log_keys=None, can_sum=True):
        _init_base(self, task)
        self.loss_weights = loss_weights
        self.log_keys = log_keys
        self.can_sum_original = can_sum
        self.can_sum = can_sum

    def forward(self, model, sample, reduce=True):
        if model.training:
            self.can_sum = self.can_sum_original
        net_output = model(**sample["net_input"])
        scaled_losses = {}
        if hasattr(model, "get_losses"):
            losses = model.get_losses(net_output, sample)
        elif isinstance(net_output, dict) and "losses" in net_output:
            losses = net_output["losses"]
        else:
            raise Exception("Could not retrieve losses")
        for lk, p in losses.items():
            try:
                coef = 1.0 if not self.loss_weights else self.loss_weights[lk]
            except KeyError:
                raise KeyError(f"weight for loss {lk} is not in loss_weights ({self.loss_weights})")
            if coef != 0 and p is not None:
                scaled_losses[lk] = coef * p.float().sum()
        loss = sum(scaled_losses.values())
        sample_size = net_output["sample_size"] if "sample_size" in net_output else loss.numel()
        if reduce and loss.numel() > 1:
            loss = loss.sum()
        logging_output = {
            "loss": loss.data,
            "ntokens": sample_size,
            "nsentences": sample["id"].numel() if "id" in sample else 0,
            "sample_size": sample_size,
            "_world_size": 1,
        }
        for lk in self.log_keys or []:
            if lk in net_output and net_output[lk] is not None:
                v = net_output[lk]
                if not torch.is_tensor(v) or v.numel() == 1:
                    logging_output[lk] = float(v)
                elif lk.startswith("_"):
                    logging_output[lk] = v
                else:
                    for i, vv in enumerate(v):
                        logging_output[f"{lk}_{i}"] = float(vv)
        if len(scaled_losses) > 1:
            for lk, l in scaled_losses.items():
                if l.numel() > 1:
                    l = l.sum()
                logging_output[f"loss_{lk}"] = l.item()
        if "logs" in net_output:
            for lgw in net_output["logs"]:
                logging_output[lgw] = net_output["logs"][lgw]
        if not model.training:
            self.can_sum = False
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs) -> Dict[str, float]:
        """nn/criterions.py:413-447 without fairseq's global ``metrics`` aggregator: returns the scalars it
        would log (loss per sample in base e, sums of the counters, world-averaged custom keys)."""
        def item(v):
            return float(v.item() if torch.is_tensor(v) else v)

        loss_sum = sum(item(l.get("loss", 0)) for l in logging_outputs)
        sample_size = sum(item(l.get("sample_size", 0)) for l in logging_outputs)
        out = {"loss": loss_sum / max(sample_size, 1.0),
               "ntokens": sum(item(l.get("ntokens", 0)) for l in logging_outputs),
               "nsentences": sum(item(l.get("nsentences", 0)) for l in logging_outputs),
               "sample_size": sample_size}
        world = sum(item(l.get("_world_size", 0)) for l in logging_outputs)
        builtin = {"loss", "ntokens", "nsentences", "sample_size", "_world_size"}
        for k in logging_outputs[0]:
            if k not in builtin and not k.startswith("_"):
                val = sum(item(l.get(k, 0)) for l in logging_outputs)
                if k.startswith("loss_"):
                    out[k] = val / max(sample_size, 1.0)
                elif k.startswith("pretrain/"):
                    out[k] = val
                else:
                    out[k] = val / max(world, 1.0)
        return out

    def logging_outputs_can_be_summed(self) -> bool:
        return self.can_sum
