"""Golden fixture for BASELINE configs[4]: animal2vec-large pretraining forward on ONE 10-s 48 kHz clip (480 000 samples,
T = 12 000 frames, ~852 kept tokens per clone). The reference itself cannot run this configuration (its ALiBi tensor
alone is 9.2 GB fp32 per copy), so the fixture comes from the CPU oracle -- pinned against the
reference at 8 kHz by tests/test_oracle.py -- with the ALiBi bias evaluated lazily (oracle.LazyAlibi, same values).
Takes several minutes of CPU time; the GPU test compares against the stored loss / masks / sampled stage outputs.

    python tests/golden/make_golden_48k.py
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import a2v_oracle as O  # noqa: E402


def main():
    cfg = O.large_config(sample_rate=48000)
    params = O.init_params(cfg, 0)
    n = 480000
    x = F.layer_norm(torch.randn(1, n, generator=torch.Generator().manual_seed(4)), (n,))
    ids = torch.arange(1) + 5
    teacher = O.make_teacher(params)
    taps = {}
    t0 = time.time()
    with torch.no_grad():
        res = O.pretrain_forward(params, teacher, cfg, x, ids, 3, taps=taps)
    loss = res["losses"]["AUDIO_regression"].double().sum()
    mask = res["mask"].numpy()
    out = {"n": np.int64(n), "seed_x": np.int64(4), "id0": np.int64(5), "num_updates": np.int64(3), "seed_w": np.int64(0),
           "T": np.int64(mask.shape[1]), "mask_packed": np.packbits(mask, axis=1), "loss_sum": np.float64(loss),
           "sample_size": np.int64(res["sample_size"]), "masked_pct": np.float64(res["masked_pct"]),
           "pred_var": np.float64(res["pred_var"]), "target_var": np.float64(res["target_var"]),
           "targets": taps["targets"][:, ::97, ::41].float().numpy(),
           "local_features": taps["local_features"][:, ::97, ::41].float().numpy(),
           "decoder_out": taps["decoder_out"][:, ::197, ::41].float().numpy()}
    path = os.path.join(HERE, "large_48k.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "loss", float(loss), "T", mask.shape, "kept", int((~mask[0]).sum()), "seconds", time.time() - t0)


if __name__ == "__main__":
    main()
