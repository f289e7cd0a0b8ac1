"""Task / dataset surface of the reference (/root/reference/nn/audio_tasks.py): registry name ``audio_ccas``,
``AudioConfigCCAS`` (:41-89 on top of fairseq's AudioPretrainingConfig), ``FileAudioLabelDataset`` (:190-469).

On-disk formats kept: ``<split>.tsv`` manifest (first line = root directory, then ``relative/path<TAB>num_samples``),
audio files under a ``wav`` directory, label files under the sibling ``lbl`` directory as ``.h5`` with the datasets
``start_frame_lbl, end_frame_lbl, lbl_cat, foc`` (nn/audio_tasks.py:336-345; ``.npz`` with the same array names is
accepted too -- the reference's own ``filename_audio2label`` defaults to it, and h5py is an optional import here).

What moves to the device: the collater hands out pinned host batches; ``to_device`` copies
them asynchronously and runs the per-clip layer norm (task.normalize) and the interval -> frame-level multi-hot target
construction as two kernels for the whole batch (``a2v_clip_layer_norm``, ``a2v_frame_labels``) instead of per clip on
the host. ``__getitem__`` / ``collater`` also provide the host results (numpy) in the reference's own format.
"""
from __future__ import annotations

import ast
import os
import re
import wave
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import registry
from .config import SHIPPED_CONV_LAYERS, parse_conv_layers


@dataclass
class AudioConfigCCAS:
    """nn/audio_tasks.py:41-89 + the fields of fairseq's AudioPretrainingConfig the dataset reads."""

    data: Optional[str] = None
    labels: Optional[str] = None
    sample_rate: int = 8000
    normalize: bool = False
    enable_padding: bool = False
    max_sample_size: Optional[int] = None
    min_sample_size: Optional[int] = None
    num_batch_buckets: int = 0
    text_compression_level: str = "none"
    do_focal_prediction: bool = True
    with_labels: bool = False
    verbose_tensorboard_logging: bool = False
    min_label_size: int = 0
    split: Optional[str] = "pretrain"
    unique_labels: Optional[str] = ("['beep', 'synch', 'eating', 'cc', 'oth', 'sn', 'ld', 'mo', 'agg', 'al', 'soc', "
                                    "'focal']")
    conv_feature_layers: str = "[(63, 125, 1)] +[(512, 10, 5)] + [(512, 3, 2)] * 3 + [(512, 3, 1)] + [(512, 2, 1)] * 2"
    train_subset: Optional[str] = None  # II("dataset.train_subset")
    valid_subset: Optional[str] = None  # II("dataset.valid_subset")
    use_focal_loss: bool = True         # II("criterion.use_focal_loss")
    segmentation_metrics: bool = False  # II("criterion.segmentation_metrics")


def feature_frames(n_samples: int, conv_layers: Sequence[Sequence[int]]) -> int:
    """Number of label frames for a clip: nn/audio_tasks.py:347-349 with nn/utils.py:80-98 (kernel capped at 10, padding
    ceil(stride / 2), stride-1 layers keep the length)."""
    size = int(n_samples)
    for (_dim, k, stride) in conv_layers:
        if stride == 1:
            continue
        p = int(np.ceil(stride / 2))
        size = int(np.floor((size + 2 * p - (min(10, k) - 1) - 1) / stride + 1))
    return size


def frame_targets(wav_len: int, frames: int, start: Sequence[int], end: Sequence[int], cat: Sequence[int],
                  foc: Optional[Sequence[int]], unique_labels: Sequence[str], use_focal_loss: bool = True,
                  do_focal_prediction: bool = True) -> np.ndarray:
    """nn/audio_tasks.py:351-381 on the host: sample-level label vector, read at
    round(linspace(0, wav_len, frames, endpoint=False)) (interp1d at integer positions = indexing)."""
    idx = np.round(np.linspace(0, wav_len, frames, endpoint=False)).astype(np.int64)
    if use_focal_loss:
        vec = np.zeros((wav_len, len(unique_labels)), dtype=np.int64)
    else:
        vec = np.zeros((wav_len,), dtype=np.int64)
    if len(start) > 0 and len(end) > 0 and len(cat) > 0:
        for ii, (s, e, l) in enumerate(zip(start, end, cat)):
            if use_focal_loss:
                vec[int(s):int(e), int(l)] = 1
                if do_focal_prediction and unique_labels[-1].lower() == "focal" and foc is not None and foc[ii] == 1:
                    vec[int(s):int(e), -1] = 1
            else:
                vec[int(s):int(e)] = int(l) + 1
    return vec[idx]


def read_audio(path: str):
    """(float32 samples in [-1, 1), sample rate). soundfile when importable (the reference's reader,
    nn/audio_tasks.py:317,330), otherwise PCM .wav through the standard library."""
    try:
        import soundfile as sf  # type: ignore

        wav, sr = sf.read(path, dtype="float32")
        return wav, sr
    except ImportError:
        pass
    with wave.open(path, "rb") as f:
        sr, nch, width, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        wav = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        wav = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        wav = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported PCM sample width {width} in {path}")
    if nch > 1:
        wav = wav.reshape(-1, nch)
    return wav, sr


def read_labels(path: str) -> Dict[str, np.ndarray]:
    keys = ("start_frame_lbl", "end_frame_lbl", "lbl_cat", "foc")
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: np.asarray(z[k]) if k in z.files else np.zeros(0, dtype=np.int64) for k in keys}
    import h5py  # type: ignore  (optional dependency, as in the reference)

    with h5py.File(path, "r") as f:
        return {k: np.asarray(f[k]) if k in f else np.zeros(0, dtype=np.int64) for k in keys}


class FileAudioLabelDataset:
    """nn/audio_tasks.py:190-469."""

    audio2label_re = re.compile(r"(?P<pre>.*)(?P<dir>wav)(?P<post>/.*\.)(?P<ext>[a-z0-9]+)$", re.IGNORECASE)

    def __init__(self, manifest_path, sample_rate, max_sample_size=None, min_sample_size=0, shuffle=True, pad=False,
                 normalize=False, unique_labels=None, use_focal_loss=None, return_labels=True,
                 conv_feature_layers=SHIPPED_CONV_LAYERS, min_label_size=0, segmentation_metrics=None,
                 do_focal_prediction=True, label_ext="h5"):
        if pad:
            raise NotImplementedError("enable_padding (the model path rejects padding masks)")
        self.sample_rate = sample_rate
        self.max_sample_size = max_sample_size if max_sample_size is not None else 2 ** 62
        self.min_sample_size = min_sample_size
        self.shuffle, self.pad, self.normalize = shuffle, pad, normalize
        self.unique_labels = list(unique_labels) if unique_labels is not None else None
        self.use_focal_loss, self.return_labels = use_focal_loss, return_labels
        self.do_focal_prediction, self.segmentation_metrics = do_focal_prediction, segmentation_metrics
        self.conv_feature_layers = parse_conv_layers(conv_feature_layers) if isinstance(conv_feature_layers, str) \
            else list(conv_feature_layers)
        self.label_ext = label_ext
        self.fnames: List[str] = []
        sizes = []
        self.skipped_indices = set()
        with open(manifest_path, "r") as f:
            self.root_dir = f.readline().strip()
            parents, last = os.path.split(self.root_dir)
            self.label_dir = parents if last in ("wav", "flac", "audio") else self.root_dir
            for i, line in enumerate(f):
                items = line.strip().split("\t")
                assert len(items) == 2, line
                sz = int(items[1])
                label_size = 0.0
                if return_labels or min_label_size > 0:
                    lf = self.filename_audio2label(os.path.join(self.root_dir, items[0]), lblext=label_ext)
                    label_size = os.path.getsize(lf) if os.path.isfile(lf) else 0.0
                    if not return_labels and min_label_size <= 0:
                        label_size = 1.0
                else:
                    label_size = 1.0
                if (min_sample_size is not None and sz < min_sample_size) or label_size <= min_label_size:
                    self.skipped_indices.add(i)
                    continue
                self.fnames.append(items[0])
                sizes.append(sz)
        self.sizes = np.array(sizes, dtype=np.int64)

    def __len__(self):
        return len(self.fnames)

    def filename_audio2label(self, audiofile, lbldir="lbl", lblext="npz"):
        m = self.audio2label_re.match(audiofile)
        if m is None:
            raise RuntimeError(f"Cannot derive label file from: {audiofile}")
        return m.expand(rf"\g<pre>{lbldir}\g<post>{lblext}")

    def postprocess(self, feats: torch.Tensor, curr_sample_rate: int, normalize: Optional[bool] = None) -> torch.Tensor:
        """fairseq RawAudioDataset.postprocess: channel mean, sample-rate check, per-clip layer norm."""
        if feats.dim() == 2:
            feats = feats.mean(-1)
        if curr_sample_rate != self.sample_rate:
            raise Exception(f"sample rate: {curr_sample_rate}, need {self.sample_rate}")
        assert feats.dim() == 1, feats.dim()
        if self.normalize if normalize is None else normalize:
            with torch.no_grad():
                feats = torch.nn.functional.layer_norm(feats, feats.shape)
        return feats

    def __getitem__(self, index, *, host_postprocess: bool = True):
        fn = self.fnames[index]
        wav, sr = read_audio(os.path.join(self.root_dir, fn))
        feats = self.postprocess(torch.from_numpy(np.ascontiguousarray(wav)).float(), sr,
                                 normalize=None if host_postprocess else False)
        item = {"id": index, "source": feats}
        if self.return_labels:
            lbl = read_labels(self.filename_audio2label(os.path.join(self.label_dir, fn), lblext=self.label_ext))
            wav_len = len(np.asarray(wav).squeeze()) if np.asarray(wav).ndim == 1 else np.asarray(wav).shape[0]
            frames = feature_frames(wav_len, self.conv_feature_layers)
            item["intervals"] = (lbl["start_frame_lbl"].astype(np.int64), lbl["end_frame_lbl"].astype(np.int64),
                                 lbl["lbl_cat"].astype(np.int64), lbl["foc"].astype(np.int64))
            item["frames"], item["wav_len"] = frames, wav_len
            if host_postprocess:
                s, e, c, fo = item["intervals"]
                item["target"] = frame_targets(wav_len, frames, s, e, c, fo if len(fo) else None, self.unique_labels,
                                               bool(self.use_focal_loss), self.do_focal_prediction)
        return item

    def collater(self, samples, *, host_postprocess: bool = True):
        """nn/audio_tasks.py:433-469 for equal-length clips (the shipped 10-s recipe): stacks sources / targets. With
        ``host_postprocess=False`` the batch carries raw audio + flat label intervals in pinned memory for
        :meth:`to_device` (layer norm and frame targets then run on the GPU)."""
        samples = [s for s in samples if s["source"] is not None]
        if not samples:
            return {}
        sizes = [len(s["source"]) for s in samples]
        target_size = min(min(sizes), self.max_sample_size)
        if any(sz != target_size for sz in sizes):
            raise NotImplementedError("clips of different length in one batch (random crops, nn/audio_tasks.py:408-421)")
        src = torch.stack([s["source"] for s in samples]).contiguous()
        out = {"id": torch.LongTensor([s["id"] for s in samples]), "net_input": {"source": src}}
        if self.return_labels:
            if host_postprocess:
                tg = torch.from_numpy(np.stack([s["target"] for s in samples]))
                out["net_input"]["target"] = tg
                out["target"] = tg
                out["ntokens"] = sum(len(t) for t in tg)
            else:
                offs = np.cumsum([0] + [len(s["intervals"][0]) for s in samples]).astype(np.int32)
                cat4 = [np.concatenate([s["intervals"][k] for s in samples]).astype(np.int32) if offs[-1] else
                        np.zeros(0, dtype=np.int32) for k in range(4)]
                if len(cat4[3]) != len(cat4[0]):  # label files without focal flags
                    cat4[3] = np.zeros(len(cat4[0]), dtype=np.int32)
                out["_intervals"] = {"offsets": torch.from_numpy(offs), "start": torch.from_numpy(cat4[0]),
                                     "end": torch.from_numpy(cat4[1]), "cat": torch.from_numpy(cat4[2]),
                                     "foc": torch.from_numpy(cat4[3]), "frames": samples[0]["frames"],
                                     "wav_len": samples[0]["wav_len"]}
                out["ntokens"] = samples[0]["frames"] * len(samples)
        if not host_postprocess:
            out["net_input"]["source"] = src.pin_memory() if torch.cuda.is_available() else src
        return out

    def to_device(self, batch: dict, device="cuda") -> dict:
        """Pinned batch -> device: asynchronous copies, then per-clip layer norm and frame-level targets as two kernels."""
        from . import ops

        src = batch["net_input"]["source"].to(device, non_blocking=True)
        if self.normalize:
            src = ops.clip_layer_norm(src)
        out = {"id": batch["id"], "net_input": {"source": src}}
        iv = batch.get("_intervals")
        if iv is not None:
            d = {k: iv[k].to(device, non_blocking=True) for k in ("offsets", "start", "end", "cat", "foc")}
            focal_class = (len(self.unique_labels) - 1) if (self.do_focal_prediction and self.unique_labels
                                                            and self.unique_labels[-1].lower() == "focal") else -1
            tg = ops.frame_labels(d["offsets"], d["start"], d["end"], d["cat"], d["foc"], src.shape[0], iv["frames"],
                                  len(self.unique_labels), iv["wav_len"], focal_class)
            out["net_input"]["target"] = tg
            out["target"] = tg
            out["ntokens"] = batch["ntokens"]
        return out


_TaskBase = registry.fairseq_bases()[2]


@registry.register_task("audio_ccas", dataclass=AudioConfigCCAS)
class AudioTaskCCAS(_TaskBase):
    """nn/audio_tasks.py:92-160."""

    def __init__(self, cfg: AudioConfigCCAS, **kwargs):
        if _TaskBase is not object:
            super().__init__(cfg, **kwargs)
        self.cfg = cfg
        self.datasets: Dict[str, FileAudioLabelDataset] = {}
        self.unique_labels = ast.literal_eval(cfg.unique_labels) if (cfg.unique_labels and cfg.with_labels) else None
        self.use_focal_loss = cfg.use_focal_loss if cfg.with_labels else None
        self.segmentation_metrics = cfg.segmentation_metrics if cfg.with_labels else None

    @classmethod
    def setup_task(cls, cfg: AudioConfigCCAS, **kwargs):
        return cls(cfg, **kwargs)

    def load_dataset(self, split: str, task_cfg: Optional[AudioConfigCCAS] = None, **kwargs):
        task_cfg = task_cfg or self.cfg
        manifest_path = os.path.join(self.cfg.data, f"{split}.tsv")
        self.datasets[split] = FileAudioLabelDataset(
            manifest_path=manifest_path, sample_rate=task_cfg.sample_rate, max_sample_size=task_cfg.max_sample_size,
            min_sample_size=task_cfg.min_sample_size, pad=task_cfg.enable_padding, normalize=task_cfg.normalize,
            return_labels=task_cfg.with_labels, unique_labels=self.unique_labels, use_focal_loss=self.use_focal_loss,
            min_label_size=task_cfg.min_label_size, conv_feature_layers=task_cfg.conv_feature_layers,
            segmentation_metrics=self.segmentation_metrics, do_focal_prediction=task_cfg.do_focal_prediction,
            label_ext=kwargs.get("label_ext", "h5"))
        return self.datasets[split]

    def dataset(self, split: str):
        return self.datasets[split]

    def build_model(self, cfg, from_checkpoint=False):
        name = getattr(cfg, "_name", None) or (cfg.get("_name") if isinstance(cfg, dict) else None)
        cls = registry.MODELS.get(name) if name else None
        if cls is None:
            raise KeyError(f"model {name!r} is not registered")
        return cls.build_model(cfg, self)

    @property
    def source_dictionary(self):
        return None

    @property
    def target_dictionary(self):
        return None
