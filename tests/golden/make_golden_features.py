"""Golden fixture for the reference's inference / finetune entry `Data2VecMultiModel.extract_features`
(nn/data2vec2.py:1112-1123 -> forward(features_only=True, mask=False), :632-728): the UNMODIFIED reference in eval mode
on CPU, weights from oracle.a2v_oracle.init_params. First fixture of the "next" row SURVEY.md section 8(f)-1 (finetune
path); the B200 implementation of that row is not part of round 1 -- the oracle and its pin come first.

    python tests/golden/make_golden_features.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import a2v_oracle as O  # noqa: E402
import make_golden as G  # noqa: E402


def main():
    cfg = O.tiny_config()
    model = G.build_reference_model(cfg, dropout=False, mixup=False, noise=False)
    params = O.init_params(cfg, 0)
    G.load_params(model, params)
    model.eval()
    b, n = 2, 16000
    x = F.layer_norm(torch.randn(b, n, generator=torch.Generator().manual_seed(5)), (n,))
    with torch.no_grad():
        res = model.extract_features(x, mode=None, padding_mask=None, mask=False)
    assert res["mask"] is None and res["padding_mask"] is None and res["linear_eval_projection"] is None
    lrs = res["layer_results"]
    out = {"b": np.int64(b), "n": np.int64(n), "seed_x": np.int64(5), "x": G.sub(res["x"]),
           "n_layers": np.int64(len(lrs)), "x_norm": np.float64(res["x"].double().norm()),
           "layer_norms": np.array([float(l.double().norm()) for l in lrs])}
    for i, l in enumerate(lrs):
        out[f"layer{i}"] = G.sub(l)
    # the finetune head's input (nn/wav2vec2.py:446-462): mean of the top-k FFN outputs
    k = cfg.average_top_k_layers
    out["topk_mean"] = G.sub(sum(lrs[-k:]) / len(lrs[-k:]))
    # finetune head + criterion pieces (nn/wav2vec2.py:463-464 proj = nn.Linear(D, classes); nn/criterions.py:231-246 focal
    # loss branch, :218-229 confusion counters) evaluated with the REFERENCE's own functions on seeded head weights / labels
    from nn.utils import confusion, sigmoid_focal_loss  # the reference's implementations

    classes = 12
    gen = torch.Generator().manual_seed(9)
    w = torch.randn(classes, cfg.embed_dim, generator=gen) * 0.2
    bias = torch.randn(classes, generator=gen) * 0.1
    top = sum(lrs[-k:]) / len(lrs[-k:])
    logits = F.linear(top, w, bias)
    target = (torch.rand(logits.shape, generator=gen) < 0.15).float()
    loss_none = sigmoid_focal_loss(logits, target, reduction="none")
    loss_sum = sigmoid_focal_loss(logits, target, reduction="sum")
    thr = 0.5
    preds = torch.where(torch.sigmoid(logits.view(-1, classes)) < thr, 0, 1)
    tp, fp, tn, fn = confusion(preds, target.view(-1, classes).to(torch.int64))
    out.update({"head_seed": np.int64(9), "classes": np.int64(classes), "logits": G.sub(logits, rows=7, cols=1),
                "focal_none": G.sub(loss_none, rows=7, cols=1), "focal_sum": np.float64(loss_sum.double()),
                "metric_threshold": np.float64(thr), "confusion": np.array([int(tp), int(fp), int(tn), int(fn)], dtype=np.int64)})
    path = os.path.join(HERE, "tiny_features.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k_: getattr(v, "shape", v) for k_, v in out.items()})


if __name__ == "__main__":
    main()
