"""Input pipeline on the device: the collated batch's per-clip layer norm and interval ->
frame-level multi-hot targets as two kernels, against the host path that is pinned to the reference's own Dataset output
(tests/test_host_cpu.py::test_dataset_frame_targets_match_the_reference_dataset)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_device_batch_matches_host_batch(tmp_path):
    from animal2vec_b200 import audio_tasks as AT
    from test_host_cpu import _write_dataset

    g = np.load(os.path.join(GOLD, "labels.npz"))
    labels = [str(x) for x in g["labels"]]
    _write_dataset(tmp_path, g)
    cfg = AT.AudioConfigCCAS(data=str(tmp_path), normalize=True, with_labels=True, unique_labels=str(labels),
                             conv_feature_layers=str(g["conv"]), min_sample_size=1000)
    ds = AT.AudioTaskCCAS.setup_task(cfg).load_dataset("train", label_ext="npz")
    host = ds.collater([ds[0], ds[1]])
    raw = ds.collater([ds.__getitem__(0, host_postprocess=False), ds.__getitem__(1, host_postprocess=False)],
                      host_postprocess=False)
    assert raw["net_input"]["source"].is_pinned() and "_intervals" in raw
    dev = ds.to_device(raw, "cuda")
    assert torch.equal(dev["target"].cpu().long(), host["target"])  # bit exact: integer work
    src = dev["net_input"]["source"].cpu()
    assert torch.allclose(src, host["net_input"]["source"], atol=2e-5, rtol=1e-5)
    assert dev["ntokens"] == host["ntokens"]


@pytest.mark.parametrize("b,n", [(3, 80000), (2, 31999), (1, 7)])
def test_clip_layer_norm_and_frame_labels_kernels(b, n):
    from animal2vec_b200 import ops
    from animal2vec_b200.audio_tasks import frame_targets

    gcpu = torch.Generator().manual_seed(b * 1000 + n)
    x = (torch.randn(b, n, generator=gcpu) * 3 + 5)
    y = ops.clip_layer_norm(x.cuda())
    assert torch.allclose(y.cpu(), F.layer_norm(x, (n,)), atol=3e-5, rtol=1e-4)
    if n < 1000:
        return
    rng = np.random.default_rng(n)
    labels = [f"c{i}" for i in range(11)] + ["focal"]
    frames = 2000 if n == 80000 else 800
    offs, st, en, ca, fo, want = [0], [], [], [], [], []
    for _ in range(b):
        k = int(rng.integers(0, 7))
        s = np.sort(rng.integers(0, n - 100, k)); e = s + rng.integers(1, 4000, k)
        c = rng.integers(0, 11, k); f = rng.integers(0, 2, k)
        want.append(frame_targets(n, frames, s, np.minimum(e, n), c, f, labels))
        offs.append(offs[-1] + k); st += list(s); en += list(np.minimum(e, n)); ca += list(c); fo += list(f)
    t = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    got = ops.frame_labels(t(offs), t(st), t(en), t(ca), t(fo), b, frames, 12, n, 11)
    assert np.array_equal(got.cpu().numpy().astype(np.int64), np.stack(want))
