"""Golden fixture for the input pipeline: the UNMODIFIED reference
`FileAudioLabelDataset.__getitem__` (nn/audio_tasks.py:316-386) -- per-clip layer norm through the restated fairseq
`RawAudioDataset.postprocess`, frame-level multi-hot targets from label intervals (scipy interp1d) -- run on three
synthetic clips. soundfile / h5py are absent here: the reference's `sf.read` and `h5py.File` calls are served by small
in-memory stand-ins holding the synthetic data (no arithmetic of the reference is replaced).

    python tests/golden/make_golden_labels.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

LABELS = ['beep', 'synch', 'sn', 'cc', 'ld', 'oth', 'mo', 'al', 'soc', 'agg', 'eating', 'focal']
CONV = "[(127, 63, 1)] +[(512, 10, 5)] + [(512, 3, 2)] * 3 + [(512, 3, 1)] + [(512, 2, 1)] * 2"


def synth(seed, n):
    g = np.random.default_rng(seed)
    wav = (g.standard_normal(n) * 0.1 + 0.02).astype(np.float32)
    k = int(g.integers(3, 9))
    start = np.sort(g.integers(0, n - 4000, k))
    end = start + g.integers(37, 3500, k)
    cat = g.integers(0, 11, k)
    foc = g.integers(0, 2, k)
    return wav, start.astype(np.int64), end.astype(np.int64), cat.astype(np.int64), foc.astype(np.int64)


def main():
    ref_shims.import_reference()
    import nn.audio_tasks as RT

    clips = {f"wav/c{i}.wav": synth(40 + i, n) for i, n in enumerate((80000, 80000, 31999))}

    class FakeSF:
        @staticmethod
        def read(path, dtype="float32"):
            key = "wav/" + os.path.basename(path)
            return clips[key][0].copy(), 8000

    class FakeH5File(dict):
        def __init__(self, path, mode="r"):
            key = "wav/" + os.path.basename(path).replace(".h5", ".wav")
            _w, s, e, c, f = clips[key]
            super().__init__(start_frame_lbl=s, end_frame_lbl=e, lbl_cat=c, foc=f)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    import types

    sf_mod = types.ModuleType("soundfile")
    sf_mod.read = FakeSF.read
    sys.modules["soundfile"] = sf_mod  # `import soundfile as sf` inside the reference's __getitem__
    RT.h5py = type("h5py", (), {"File": FakeH5File})
    RT.parse_path = lambda p: (p, [])

    ds = object.__new__(RT.FileAudioLabelDataset)
    ds.fnames = list(clips.keys())
    ds.text_compressor = type("TC", (), {"decompress": staticmethod(lambda x: x)})()
    ds.root_dir = "/data/root"
    ds.label_dir = "/data/root"
    ds.return_labels, ds.use_focal_loss, ds.do_focal_prediction = True, True, True
    ds.unique_labels = LABELS
    ds.conv_feature_layers = eval(CONV)
    ds.sample_rate, ds.normalize = 8000, True

    def postprocess(feats, curr_sample_rate):  # fairseq RawAudioDataset.postprocess (third party, restated)
        if feats.dim() == 2:
            feats = feats.mean(-1)
        assert curr_sample_rate == ds.sample_rate and feats.dim() == 1
        with torch.no_grad():
            feats = torch.nn.functional.layer_norm(feats, feats.shape)
        return feats

    ds.postprocess = postprocess
    out = {"n_clips": np.int64(len(clips)), "labels": np.array(LABELS), "conv": np.array(CONV)}
    for i, key in enumerate(clips):
        item = ds[i]
        wav, s, e, c, f = clips[key]
        tg = np.asarray(item["target"])
        out[f"seed{i}"], out[f"n{i}"] = np.int64(40 + i), np.int64(len(wav))  # regenerate with synth(seed, n)
        out[f"target{i}"] = np.packbits(tg.astype(bool), axis=1)
        out[f"target_shape{i}"] = np.array(tg.shape, dtype=np.int64)
        out[f"source_head{i}"] = item["source"][:64].numpy()
        out[f"source_sum{i}"] = np.float64(item["source"].double().sum())
        out[f"source_sq{i}"] = np.float64(item["source"].double().pow(2).sum())
        print(key, "target", tg.shape, "positives", int(tg.sum()))
    path = os.path.join(HERE, "labels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
