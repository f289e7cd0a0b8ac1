"""Host-side mask-index generation for the multi-mask clone batch (integer work, bit exact).

Follows the reference call chain nn/data2vec2.py:618-620 (MaskSeed) -> nn/modalities/base.py:246-259
(per-clone seed ids) -> base.py:370-425 (compute_mask) -> fairseq ``compute_mask_indices`` @ 920a548
(third party, not vendored by the reference; behaviour restated in SURVEY.md Appendix B1): one
``numpy.random.default_rng`` (PCG64) per row seeded with ``int(hash((seed, update, id)) % 1e6)``,
``Generator.choice(replace=False)`` for the span starts, spans of ``mask_length``, and an equalisation
of every row to the batch-minimum masked count that re-uses the LAST row's generator.

The masks depend only on ``(seed, update, ids)``, so :class:`MaskPrefetcher` computes step n+1's masks
on a worker thread while step n runs on the GPU (the reference computes them inline: a host stall).
Python's tuple-of-int ``hash`` is interpreter-version dependent; reference and replacement must run
on the same interpreter for seeded masks to agree.
"""
from __future__ import annotations

import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np


def clone_seed_ids(seed: int, ids: np.ndarray, clone_batch: int) -> np.ndarray:
    """base.py:246-259: id of clone m of clip i = ids[i] + clone_hash[m], clone_hash[0] = 0."""
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    if clone_batch <= 1:
        return ids
    clone_hash = np.array([0] + [int(hash((seed, ind)) % 1e10) for ind in range(clone_batch - 1)], dtype=np.int64)
    return (ids[:, None] + clone_hash[None, :]).reshape(-1)


def compute_mask_indices(bsz: int, all_sz: int, mask_prob: float, mask_length: int, *, seed: Optional[int],
                         epoch: Optional[int], indices: Optional[np.ndarray], min_masks: int = 0,
                         require_same_masks: bool = True, mask_dropout: float = 0.0,
                         add_masks: bool = False) -> np.ndarray:
    """Static-length, overlap-allowed branch of fairseq's compute_mask_indices (the only branch the
    shipped configs reach through base.py:401-413). Returns a (bsz, all_sz) bool array."""
    mask = np.zeros((bsz, all_sz), dtype=bool)
    rows = []
    rng = None
    offs = np.arange(mask_length, dtype=np.int64)
    for i in range(bsz):
        if seed is not None and epoch is not None and indices is not None:
            seed_i = int(hash((seed, epoch, int(indices[i]))) % 1e6)
        else:
            seed_i = None  # OS entropy, as upstream
        rng = np.random.default_rng(seed_i)
        sz = all_sz
        num_mask = max(min_masks, int(mask_prob * sz / float(mask_length) + rng.random()))
        min_len = mask_length
        if sz - min_len <= num_mask:
            min_len = sz - num_mask - 1
        starts = rng.choice(sz - min_len, num_mask, replace=False)
        idc = (starts.astype(np.int64)[:, None] + offs[None, :]).reshape(-1)
        idc = np.unique(idc[idc < sz])
        if len(idc) >= sz:
            raise ValueError(f"the entire sequence is masked. sz={sz}; mask_idc[mask_idc]; index={i}")
        rows.append(idc)
    target_len = None
    if require_same_masks:
        target_len = max(len(m) for m in rows) if add_masks else min(len(m) for m in rows)
    for i, idc in enumerate(rows):
        if target_len is not None and len(idc) > target_len:
            idc = rng.choice(idc, target_len, replace=False)
        mask[i, idc] = True
        if target_len is not None and len(idc) < target_len:
            unmasked = np.flatnonzero(~mask[i])
            mask[i, rng.choice(unmasked, target_len - len(idc), replace=False)] = True
        if mask_dropout > 0:
            masked = np.flatnonzero(mask[i])
            num_holes = np.rint(len(masked) * mask_dropout).astype(int)
            mask[i, rng.choice(masked, num_holes, replace=False)] = False
    return mask


def pretrain_mask(*, seed: int, update: int, ids: Optional[Sequence[int]], batch: int, frames: int, clone_batch: int,
                  mask_prob: float, mask_length: int, mask_dropout: float = 0.0, add_masks: bool = False,
                  inverse_mask: bool = False) -> np.ndarray:
    """(batch*clone_batch, frames) bool mask of one forward (True = masked)."""
    if inverse_mask:
        raise NotImplementedError("inverse_mask=True (image-style block masking) is not on the audio pretraining path")
    rows = batch * clone_batch
    idx = None
    if ids is not None:
        idx = clone_seed_ids(seed, np.asarray(ids), clone_batch)
    return compute_mask_indices(rows, frames, mask_prob, mask_length, seed=seed if ids is not None else None,
                                epoch=update if ids is not None else None, indices=idx, min_masks=1,
                                require_same_masks=True, mask_dropout=mask_dropout, add_masks=add_masks)


class MaskPrefetcher:
    """Computes masks for announced (update, ids) pairs on a worker thread."""

    def __init__(self, **static):
        self._static = static
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="a2v-mask")
        self._pending: dict = {}
        self._lock = threading.Lock()

    def announce(self, update: int, ids: Sequence[int]) -> None:
        key = (int(update), tuple(int(i) for i in ids))
        with self._lock:
            if key not in self._pending:
                self._pending[key] = self._pool.submit(pretrain_mask, update=key[0], ids=key[1], **self._static)

    def get(self, update: int, ids: Sequence[int]) -> np.ndarray:
        key = (int(update), tuple(int(i) for i in ids))
        with self._lock:
            fut: Optional[Future] = self._pending.pop(key, None)
        if fut is None:
            return pretrain_mask(update=key[0], ids=key[1], **self._static)
        return fut.result()

    def close(self) -> None:
        self._pool.shutdown(wait=False, cancel_futures=True)
