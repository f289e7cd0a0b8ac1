"""Minimal data-parallel step driver for the pretraining path.

Stands where fairseq's ``Trainer.train_step`` stands in the reference call stack
(/root/reference/nn/audio_train_routine.py:330-334 -> fairseq Trainer, third party; behaviour restated in
SURVEY.md Appendix B4): for each micro-batch backward on the SUMMED loss, all-reduce(SUM) of the
gradients, multiply by 1 / sum(sample_size), clip the global norm (``clip_norm: 1``), fairseq-style Adam
with decoupled weight decay and the ``weight_decay_scale: 0`` group of nn/data2vec2.py:318-322, cosine
schedule with linear warm-up, then ``model.set_num_updates`` (EMA teacher step).

Differences from the reference's ``legacy_ddp`` (flat all-reduce after backward, no overlap; SURVEY.md
section 2a): the gradients of every transformer block are all-reduced by NCCL on a side stream as soon
as that block's backward has finished, overlapping the rest of the backward; the six blocking
``compute_var`` all-reduces of nn/data2vec2.py:1098-1105 and the trainer's logging all-reduce are packed
into one small all-reduce per step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops
from .engine import ENC, PretrainEngine
from .params import no_decay


@dataclass
class OptimConfig:
    """configs/MeerKAT/a2v_large_pretrain_best.yaml:61-81."""

    lr: float = 1e-4
    betas: Tuple[float, float] = (0.9, 0.98)
    eps: float = 1e-6
    weight_decay: float = 0.01
    clip_norm: float = 1.0
    warmup_updates: int = 10000
    warmup_init_lr: float = 0.0
    min_lr: float = 0.0
    max_update: int = 384230


def cosine_lr(cfg: OptimConfig, num_updates: int) -> float:
    """fairseq ``cosine`` scheduler, single period: linear warm-up then half-cosine down to min_lr."""
    if num_updates < cfg.warmup_updates:
        return cfg.warmup_init_lr + (cfg.lr - cfg.warmup_init_lr) * num_updates / max(1, cfg.warmup_updates)
    period = max(1, cfg.max_update - cfg.warmup_updates)
    t = min(num_updates - cfg.warmup_updates, period)
    return cfg.min_lr + 0.5 * (cfg.lr - cfg.min_lr) * (1 + math.cos(math.pi * t / period))


class BucketReducer:
    """SUM all-reduce of contiguous slices of one flat gradient buffer, launched as the slices become
    final. On CUDA the collectives run on a dedicated stream ordered after the producing kernels by an
    event; ``finish`` makes the compute stream wait for all of them."""

    def __init__(self, flat: torch.Tensor, group=None, compress_bf16: bool = False):
        """``compress_bf16`` (CUDA): every bucket travels as bf16 -- half the NCCL bytes and half the time the
        all-reduce kernels hold SMs next to the persistent GEMMs of the ongoing backward (the measured limiter of the
        1 -> 8 GPU scaling). The fp32 buffer keeps the widened sum; rounding each rank's bucket to
        bf16 before the sum is the usual gradient-compression trade (relative error 2^-9 per addend)."""
        self.flat = flat
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.cuda = flat.is_cuda
        self.stream = torch.cuda.Stream(device=flat.device) if (self.cuda and self.enabled) else None
        self.pending: List = []
        self.bytes_reduced = 0
        self.compress = bool(compress_bf16 and self.cuda and self.enabled)
        self.lp = torch.empty(flat.numel(), device=flat.device, dtype=torch.bfloat16) if self.compress else None

    def reduce_range(self, lo: int, hi: int) -> None:
        if not self.enabled or hi <= lo:
            return
        chunk = self.flat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                if self.compress and lo % 8 == 0 and (hi - lo) % 8 == 0:
                    lp = self.lp[lo:hi]
                    ops.cast_bf16(chunk, out=lp)
                    dist.all_reduce(lp, op=dist.ReduceOp.SUM, group=self.group)
                    ops.cast_f32(lp, chunk)
                    self.bytes_reduced += lp.numel() * 2
                else:
                    dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
                    self.bytes_reduced += chunk.numel() * 4
        else:
            self.bytes_reduced += chunk.numel() * chunk.element_size()
            self.pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        if not self.enabled:
            return
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        for w in self.pending:
            w.wait()
        self.pending.clear()


class PretrainTrainer:
    def __init__(self, engine: PretrainEngine, optim: Optional[OptimConfig] = None, group=None,
                 compress_grads: Optional[bool] = None):
        """``compress_grads``: bf16 gradient buckets on the wire (default: on in bf16 mode, off in the fp32 validation
        mode; A2V_GRAD_BF16=0/1 overrides)."""
        import os as _os

        if compress_grads is None:
            env = _os.environ.get("A2V_GRAD_BF16")
            compress_grads = (not engine.fp32) if env is None else env != "0"
        self.e = engine
        self.o = optim or OptimConfig()
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        S = engine.S
        dev = engine.device
        self.m = torch.zeros_like(S.data)
        self.v = torch.zeros_like(S.data)
        mask = torch.ones(S.total // 4, dtype=torch.uint8)
        for n in S.names:
            if no_decay(n, S.shapes[n]):
                o = S.offsets[n]
                mask[o // 4:(o + S.numel(n) + 3) // 4] = 0
        self.wd_mask = mask.to(dev)
        self.num_updates = 0
        self.reducer = BucketReducer(S.grad, group, compress_bf16=compress_grads)
        # gradient buckets: one per transformer block (final when its backward ends), the rest at the end
        self.block_ranges = [S.range_of([n for n in S.names if n.startswith(pre)]) for pre in engine.block_prefixes]
        covered = sorted(self.block_ranges)
        self.tail_ranges: List[Tuple[int, int]] = []
        pos = 0
        for lo, hi in covered:
            if lo > pos:
                self.tail_ranges.append((pos, lo))
            pos = max(pos, hi)
        if pos < S.total:
            self.tail_ranges.append((pos, S.total))
        self.stats = torch.zeros(2 + 4 * engine.D, device=dev, dtype=torch.float64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.denom = torch.zeros(1, device=dev, dtype=torch.float32)
        self.coef = torch.zeros(2, device=dev, dtype=torch.float32)
        self.last_log: Dict[str, float] = {}

    def _block_done(self, j: int) -> None:
        lo, hi = self.block_ranges[j]
        self.reducer.reduce_range(lo, hi)

    def accumulate_and_reduce(self, micro_batches: Sequence[Tuple[torch.Tensor, Optional[Sequence[int]]]]):
        """Forward + backward of every micro-batch into the flat gradient buffer, bucketed SUM all-reduce of the
        gradients (overlapped with the last backward) and ONE packed all-reduce of the step statistics
        [loss sum, sample size, 4 x D column sums]. Afterwards every rank holds the global sums."""
        e = self.e
        e.zero_grad()
        self.stats.zero_()
        n_mb = len(micro_batches)
        sample_size = 0
        res = None
        for i, (source, ids) in enumerate(micro_batches):
            res = e.forward(source, ids, self.num_updates, training=True, fuse_loss_grad=True)
            last = i == n_mb - 1
            e.backward(None, training=True, block_done=self._block_done if last else None)
            self.stats[0:1] += res["loss_sum"]
            self.stats[2:] += res["colstats"].view(-1)
            sample_size += res["sample_size"]
        self.stats[1:2].fill_(float(sample_size))  # (item assignment of a Python float synchronises the stream)
        for lo, hi in self.tail_ranges:
            self.reducer.reduce_range(lo, hi)
        if self.reducer.enabled:
            dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=self.group)  # C1 + C3 packed
        self.reducer.finish()
        return res

    def train_step(self, micro_batches: Sequence[Tuple[torch.Tensor, Optional[Sequence[int]]]],
                   sync_log: bool = False) -> Dict[str, object]:
        """One optimizer update from ``len(micro_batches)`` accumulated micro-batches (``update_freq``)."""
        e, o = self.e, self.o
        res = self.accumulate_and_reduce(micro_batches)
        # grads *= 1/sum(sample_size); clip to clip_norm; Adam; EMA
        self.sumsq.zero_()
        ops.sumsq(e.S.grad, self.sumsq)
        self.denom.copy_(self.stats[1:2])
        ops.clip_coef(self.sumsq, self.denom, 1.0, float(o.clip_norm), self.coef)
        lr = cosine_lr(o, self.num_updates)
        self.num_updates += 1
        ops.adamw_step(e.S.data, e.S.grad, self.m, self.v, None if e.fp32 else e.S16, lr=lr, beta1=o.betas[0],
                       beta2=o.betas[1], eps=o.eps, weight_decay=o.weight_decay, step=self.num_updates,
                       grad_scale=self.coef[0:1], wd_mask=self.wd_mask)
        e.mark_student_updated(s16_valid=not e.fp32)
        decay = e.ema_step(self.num_updates)
        out = {"lr": lr, "ema_decay": decay * 1000, "num_updates": self.num_updates, "stats": self.stats,
               "coef": self.coef, "masked_pct": res["masked_pct"]}
        if sync_log:
            out.update(self.log_values())
        return out

    def log_values(self) -> Dict[str, float]:
        """Host copies of the step statistics (one D2H sync): loss per masked token, grad norm, variances."""
        st = self.stats.cpu()
        n = float(st[1])
        d = self.e.D
        pv, tv = PretrainEngine.variances(st[2:].view(4, d), n)
        coef = self.coef.cpu()
        self.last_log = {"loss": float(st[0]) / max(n, 1.0) / math.log(2), "loss_sum": float(st[0]), "sample_size": n,
                         "pred_var": float(pv), "target_var": float(tv), "gnorm": float(coef[1])}
        return self.last_log
