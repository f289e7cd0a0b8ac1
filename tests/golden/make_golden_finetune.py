"""Golden fixture for the finetune TRAINING step: the UNMODIFIED reference
`Data2VecMultiModel.forward(features_only=True, mask=True)` in train mode after `remove_pretraining_modules`, with the
finetune wrapper's overrides that change arithmetic (encoder_zero_mask False, channel masking, frozen feature extractor;
nn/wav2vec2.py:95-130), the head / focal loss of nn/wav2vec2.py:446-464 + nn/criterions.py:231-246 evaluated with the
reference's own functions, and its autograd gradients. Dropouts, layerdrop and mask-token noise are 0 (deterministic
stage parity); the time mask is passed as `precomputed_mask`, the channel mask through a patched
`compute_mask_indices` (the reference draws it from OS entropy).

    python tests/golden/make_golden_finetune.py
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

GRAD_KEYS = ["blocks.1.mlp.fc1.weight", "blocks.0.attn.qkv.weight", "blocks.1.mlp.fc2.bias", "blocks.0.norm2.weight",
             O.ENC + "context_encoder.blocks.0.attn.qkv.weight", O.ENC + "context_encoder.norm.weight",
             O.ENC + "alibi_scale", O.ENC + "relative_positional_encoder.1.0.weight",
             O.ENC + "relative_positional_encoder.2.0.bias"]


def masks(b, t, d):
    g = np.random.default_rng(21)
    tm = np.zeros((b, t), dtype=bool)
    for r in range(b):
        for s in g.choice(t - 4, 12, replace=False):
            tm[r, s:s + 4] = True
    cm = np.zeros((b, d), dtype=bool)
    for r in range(b):
        s = int(g.integers(0, d - 16))
        cm[r, s:s + 16] = True
    return tm, cm


def main():
    cfg = O.tiny_config()
    model = G.build_reference_model(cfg, dropout=False, mixup=False, noise=False)
    params = O.init_params(cfg, 0)
    G.load_params(model, params)
    model.remove_pretraining_modules(modality="audio")
    enc = model.modality_encoders["AUDIO"]
    mc = enc.modality_cfg
    mc.encoder_zero_mask = False       # "zero_mask": False
    mc.mask_noise_std = 0.0            # deterministic mask tokens
    mc.mask_channel_prob = 0.5
    mc.mask_channel_length = 16
    enc.local_grad_mult = 0.0          # "local_grad_mult": cfg.feature_grad_mult = 0.0
    model.train()
    b, n = 2, 16000
    x = F.layer_norm(torch.randn(b, n, generator=torch.Generator().manual_seed(6)), (n,))
    t = 400
    tm, cm = masks(b, t, cfg.embed_dim)

    import nn.modalities.base as RB
    orig = RB.compute_mask_indices

    def fixed_channel_mask(shape, *a, **k):
        assert tuple(shape) == cm.shape, shape
        return cm.copy()

    RB.compute_mask_indices = fixed_channel_mask
    try:
        res = model(x, mode=None, mask=True, features_only=True, precomputed_mask=torch.from_numpy(tm))
    finally:
        RB.compute_mask_indices = orig
    lrs = res["layer_results"]
    assert lrs[0].shape == (b, t, cfg.embed_dim), lrs[0].shape
    from nn.utils import confusion, sigmoid_focal_loss  # the reference's implementations

    classes, k = 12, cfg.average_top_k_layers
    gen = torch.Generator().manual_seed(9)
    w = (torch.randn(classes, cfg.embed_dim, generator=gen) * 0.2).requires_grad_(True)
    bias = (torch.randn(classes, generator=gen) * 0.1).requires_grad_(True)
    top = sum(lrs[-k:]) / len(lrs[-k:])
    logits = F.linear(top, w, bias)
    target = (torch.rand(logits.shape, generator=gen) < 0.15).float()
    loss = sigmoid_focal_loss(logits, target, reduction="sum")
    loss.backward()
    named = dict(model.named_parameters())
    thr = 0.5
    preds = torch.where(torch.sigmoid(logits.detach().view(-1, classes)) < thr, 0, 1)
    tp, fp, tn, fn = confusion(preds, target.view(-1, classes).to(torch.int64))
    n_correct = int(preds.eq(target.view(-1, classes).to(torch.int64)).sum())
    out = {"b": np.int64(b), "n": np.int64(n), "seed_x": np.int64(6), "head_seed": np.int64(9), "classes": np.int64(classes),
           "time_mask": np.packbits(tm, axis=1), "channel_mask": np.packbits(cm, axis=1), "T": np.int64(t),
           "logits": G.sub(logits, rows=7, cols=1), "loss_sum": np.float64(loss.double()),
           "metric_threshold": np.float64(thr),
           "confusion": np.array([int(tp), int(fp), int(tn), int(fn), n_correct], dtype=np.int64),
           "layer_last": G.sub(lrs[-1]), "grad_keys": np.array(GRAD_KEYS),
           "grad_norms": np.array([float(named[k_].grad.double().norm()) for k_ in GRAD_KEYS]),
           "grad_heads": np.stack([named[k_].grad.reshape(-1)[:8].float().numpy() if named[k_].grad.numel() >= 8
                                   else np.resize(named[k_].grad.reshape(-1).float().numpy(), 8) for k_ in GRAD_KEYS]),
           "head_w_grad_norm": np.float64(w.grad.double().norm()), "head_b_grad": bias.grad.float().numpy(),
           "fe_grad_is_none": np.bool_(named[O.ENC + "local_encoder.conv_layers.1.0.weight"].grad is None),
           "proj_feat_grad_norm": np.float64(named[O.ENC + "project_features.2.weight"].grad.double().norm()
                                             if named[O.ENC + "project_features.2.weight"].grad is not None else -1.0)}
    path = os.path.join(HERE, "tiny_finetune.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k_: getattr(v, "shape", v) for k_, v in out.items()})
    print("loss", float(loss), "confusion", out["confusion"], "fe_grad_is_none", out["fe_grad_is_none"],
          "proj_feat_grad_norm", out["proj_feat_grad_norm"])


if __name__ == "__main__":
    main()
