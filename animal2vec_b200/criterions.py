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


# ------------------------------------------------------------------------------------------------------------
# finetune criterion
# ------------------------------------------------------------------------------------------------------------
@dataclass
class FinetuneCriterionConfig:
    """nn/criterions.py:24-80 (LabelSmoothedCrossEntropyCriterionConfigModifiedLogs: fairseq's label-smoothed CE
    config + the fields of ExpandedModelCriterionConfig that the finetune criterion reads)."""

    label_smoothing: float = 0.0
    report_accuracy: bool = False
    ignore_prefix_size: int = 0
    sentence_avg: bool = False
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


class _FocalLossSum(torch.autograd.Function):
    """sigmoid_focal_loss(reduction="sum") (nn/utils.py:971-1010) on the fused kernels."""

    @staticmethod
    def forward(ctx, logits, target):
        from . import ops

        lg = logits.detach().float().contiguous().view(-1, logits.shape[-1])
        tg = target.detach().float().contiguous().view(-1, logits.shape[-1])
        loss_sum, _, _, _ = ops.focal_loss_fwd(lg, tg)
        ctx.save_for_backward(lg, tg)
        ctx.shape = logits.shape
        return loss_sum.to(torch.float32).view(())

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops

        lg, tg = ctx.saved_tensors
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        return ops.focal_loss_bwd(lg, tg, grad_out=go).view(ctx.shape), None


@registry.register_criterion("finetunecriterion", dataclass=FinetuneCriterionConfig)
class FinetuneCrossEntropyCriterion(_CriterionBase):
    """nn/criterions.py:137-277, focal-loss branch (``use_focal_loss: true`` in every shipped finetune recipe); the
    label-smoothed cross-entropy branch belongs to fairseq's LabelSmoothedCrossEntropyCriterion and is not restated."""

    def __init__(self, task, sentence_avg=False, label_smoothing=0.0, unique_labels=None, segmentation_metrics=False,
                 use_focal_loss=True, metric_threshold=0.25, iou_threshold=0.0, sigma_s=0.1, maxfilt_s=0.1,
                 max_duration_s=0.5, lowP=0.125, method="avg", can_sum=True, verbose_tensorboard_logging=False,
                 ignore_prefix_size=0, report_accuracy=False):
        _init_base(self, task)
        if not use_focal_loss:
            raise NotImplementedError("use_focal_loss=False (fairseq's label-smoothed cross entropy branch)")
        import ast

        self.sentence_avg = sentence_avg
        self.num_classes = len(ast.literal_eval(unique_labels)) if unique_labels else None
        self.segmentation_metrics = segmentation_metrics
        self.use_focal_loss = use_focal_loss
        self.metric_threshold = metric_threshold
        self.report_accuracy = report_accuracy
        self.can_sum = self.can_sum_original = can_sum
        self.verbose_tensorboard_logging = verbose_tensorboard_logging

    def forward(self, model, sample, reduce=True):
        from . import ops

        if model.training:
            self.can_sum = self.can_sum_original
        net_output = model(**sample["net_input"])
        logits = model.get_logits(net_output)
        target = model.get_targets(sample, net_output)
        if not reduce:
            raise NotImplementedError("reduce=False (unreduced focal loss) is not used by the trainer")
        loss = _FocalLossSum.apply(logits, target)
        sample_size = sample["target"].size(0) if self.sentence_avg else sample["ntokens"]
        logging_output = {"loss": loss.data, "nll_loss": torch.tensor(0.0), "ntokens": sample["ntokens"],
                          "nsentences": sample["target"].size(0), "sample_size": sample_size}
        if self.report_accuracy:
            lg = logits.detach().float().contiguous()
            tg = target.detach().float().contiguous().view(-1, lg.shape[-1])
            _, counters, _, _ = ops.focal_loss_fwd(lg, tg, threshold=float(self.metric_threshold))
            tp, fp, tn, fn, n_correct = [int(v) for v in counters.tolist()]
            logging_output.update({"finetune/n_correct": n_correct, "finetune/total": lg.numel(), "finetune/tp": tp,
                                   "finetune/fp": fp, "finetune/tn": tn, "finetune/fn": fn})
        if self.verbose_tensorboard_logging and not model.training:
            self.can_sum = False
            logging_output["_predictions"] = torch.sigmoid(model.get_logits(net_output, reshape=False))
            logging_output["_targets"] = model.get_targets(sample, net_output, reshape=False)
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs) -> Dict[str, float]:
        """nn/criterions.py:279-372 without fairseq's metrics aggregator: the scalars it would log."""
        import math

        def item(v):
            return float(v.item() if torch.is_tensor(v) else v)

        loss_sum = sum(item(l.get("loss", 0)) for l in logging_outputs)
        sample_size = sum(item(l.get("sample_size", 0)) for l in logging_outputs)
        out = {"loss": loss_sum / max(sample_size, 1.0) / math.log(2),
               "misc/ntokens": sum(item(l.get("ntokens", 0)) for l in logging_outputs),
               "misc/nsentences": sum(item(l.get("nsentences", 0)) for l in logging_outputs),
               "misc/sample_size": sample_size}
        s = {k: sum(item(l.get("finetune/" + k, 0)) for l in logging_outputs) for k in ("n_correct", "total", "tp", "fp", "tn", "fn")}
        if s["total"] > 0:
            out["metrics/finetune/accuracy"] = round(s["n_correct"] * 100.0 / s["total"], 3)
            if s["tp"] + s["fp"] > 0:
                out["metrics/finetune/precision"] = round(s["tp"] * 100.0 / (s["tp"] + s["fp"]), 3)
            if s["tp"] + s["fn"] > 0:
                out["metrics/finetune/recall"] = round(s["tp"] * 100.0 / (s["tp"] + s["fn"]), 3)
            if 2 * s["tp"] + s["fn"] + s["fp"] > 0:
                out["metrics/finetune/f1"] = round(s["tp"] * 200.0 / (2 * s["tp"] + s["fn"] + s["fp"]), 3)
        return out

    def logging_outputs_can_be_summed(self) -> bool:
        return self.can_sum
