"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import layer that lets the *unmodified* reference sources under /root/reference be imported and
run on CPU in this container (Python 3.12, torch 2.11) even though fairseq @ 920a548, timm 0.6.12,
omegaconf, hydra, tensorflow, ... are not installed. Used by ``tests/golden/make_golden.py`` to
produce the golden vectors that pin ``oracle/a2v_oracle.py`` and, through it, the CUDA path.

Two kinds of stand-ins:
  * inert stubs for packages the hot path never executes (matplotlib, tensorflow, h5py, ...);
  * behavioural restatements of the third-party pieces that DO carry arithmetic on the path
    (fairseq compute_mask_indices / EMAModule / Fp32LayerNorm / ..., timm Mlp). Those sources are
    absent from /root/reference and cannot be fetched, so parity for them is anchored on the
    reference's own call sites and documented expectations ("parity unpinned" upstream).
"""
from __future__ import annotations

import dataclasses
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import torch
import torch.nn as tnn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"

_STUB_ROOTS = ("fairseq", "omegaconf", "hydra", "timm", "matplotlib", "tensorflow", "skimage", "intervaltree",
               "h5py", "soundfile", "librosa")


# --------------------------------------------------------------------------- inert stubs
class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything(metaclass=_AnyMeta):
    """Base class for auto-generated attributes: subclassable, callable, decorator-friendly."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _identity_decorator_factory(*_a, **_k):
    def deco(obj):
        return obj

    return deco


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.startswith("register_"):
            return _identity_decorator_factory
        cls = type(name, (_Anything,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


# --------------------------------------------------------------------------- behavioural restatements
class TransposeLast(tnn.Module):
    def __init__(self, deconstruct_idx=None, tranpose_dim=-2):
        super().__init__()
        self.deconstruct_idx = deconstruct_idx
        self.tranpose_dim = tranpose_dim

    def forward(self, x):
        if self.deconstruct_idx is not None:
            x = x[self.deconstruct_idx]
        return x.transpose(self.tranpose_dim, -1)


class SamePad(tnn.Module):
    def __init__(self, kernel_size, causal=False):
        super().__init__()
        self.remove = kernel_size - 1 if causal else (1 if kernel_size % 2 == 0 else 0)

    def forward(self, x):
        if self.remove > 0:
            x = x[:, :, : -self.remove]
        return x


class Fp32LayerNorm(tnn.LayerNorm):
    def forward(self, input):
        out = F.layer_norm(input.float(), self.normalized_shape,
                           self.weight.float() if self.weight is not None else None,
                           self.bias.float() if self.bias is not None else None, self.eps)
        return out.type_as(input)


class Fp32GroupNorm(tnn.GroupNorm):
    def forward(self, input):
        out = F.group_norm(input.float(), self.num_groups,
                           self.weight.float() if self.weight is not None else None,
                           self.bias.float() if self.bias is not None else None, self.eps)
        return out.type_as(input)


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True, export=False):
    return tnn.LayerNorm(normalized_shape, eps, elementwise_affine)


class _GradMultiplyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.new(x)

    @staticmethod
    def backward(ctx, grad):
        return grad * ctx.scale, None


class GradMultiply:
    apply = _GradMultiplyFn.apply


def index_put(tensor, indices, value):
    tensor[indices] = value
    return tensor


def init_bert_params(module):
    if isinstance(module, tnn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, tnn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()


class BaseFairseqModel(tnn.Module):
    def set_num_updates(self, num_updates):
        for m in self.modules():
            if hasattr(m, "set_num_updates") and m != self:
                m.set_num_updates(num_updates)

    @classmethod
    def build_model(cls, cfg, task=None):
        raise NotImplementedError


@dataclasses.dataclass
class EMAModuleConfig:
    ema_decay: float = 0.9999
    ema_fp32: bool = False
    add_missing_params: bool = True
    log_norms: bool = False


class EMAModule:
    """fairseq.modules.EMAModule."""

    def __init__(self, model, config, copy_model=True, device=None, skip_keys=None):
        import copy

        self.model = copy.deepcopy(model) if copy_model else model
        self.model.requires_grad_(False)
        self.config = config
        self.decay = config.ema_decay
        self.skip_keys = skip_keys or set()
        self.add_missing_params = config.add_missing_params
        self.fp32_params = {}
        if device is not None:
            self.model = self.model.to(device=device)
        if self.config.ema_fp32:
            self.build_fp32_params()
        self.log_norms = False  # needs apex amp_C.multi_tensor_l2norm, absent in practice
        self.logs = {}

    def build_fp32_params(self, state_dict=None):
        if not self.config.ema_fp32:
            raise RuntimeError("build_fp32_params should not be called if ema_fp32=False")
        if state_dict is None:
            state_dict = self.model.state_dict()

        def _to_float(t):
            return t.float() if torch.is_floating_point(t) else t

        for k in state_dict:
            if k in self.fp32_params:
                if k == "__sq_mom":
                    self.fp32_params[k] = state_dict[k]
                else:
                    self.fp32_params[k].copy_(state_dict[k])
            else:
                self.fp32_params[k] = _to_float(state_dict[k])
                if "__sq_mom" in self.fp32_params:
                    self.fp32_params["__sq_mom"][k] = torch.zeros_like(self.fp32_params[k])

    def restore(self, state_dict, build_fp32_params=False):
        self.model.load_state_dict(state_dict, strict=False)
        if build_fp32_params:
            self.build_fp32_params(state_dict)

    def set_decay(self, decay, weight_decay=None):
        self.decay = decay
        if weight_decay is not None:
            self.weight_decay = weight_decay

    def get_decay(self):
        return self.decay

    @torch.no_grad()
    def step(self, new_model):
        decay = self.decay
        ema_state_dict = {}
        ema_params = self.fp32_params if self.config.ema_fp32 else self.model.state_dict()
        for key, param in new_model.named_parameters():
            if isinstance(param, dict):
                continue
            if not self.add_missing_params and key not in ema_params:
                continue
            try:
                ema_param = ema_params[key]
            except KeyError:
                ema_param = param.float().clone() if param.ndim == 1 else __import__("copy").deepcopy(param)
                ema_params[key] = ema_param
            if param.shape != ema_param.shape:
                raise ValueError("incompatible tensor shapes between model param and ema param")
            if "version" in key:
                continue
            if key in self.skip_keys or not param.requires_grad:
                ema_params[key].copy_(param.to(dtype=ema_param.dtype).data)
                ema_param = ema_params[key]
            else:
                ema_param.mul_(decay)
                ema_param.add_(param.data.to(dtype=ema_param.dtype), alpha=1 - decay)
            ema_state_dict[key] = ema_param
        for key, param in new_model.named_buffers():
            ema_state_dict[key] = param
        self.restore(ema_state_dict, build_fp32_params=False)


def compute_mask_indices(shape, padding_mask, mask_prob, mask_length, mask_type="static", mask_other=0.0,
                         min_masks=0, no_overlap=False, min_space=0, require_same_masks=True, mask_dropout=0.0,
                         add_masks=False, seed=None, epoch=None, indices=None, idc_select_ver=1, num_mask_ver=2):
    """fairseq.data.data_utils.compute_mask_indices, static-length / overlapping branch
     -- the only branch the shipped configs reach."""
    bsz, all_sz = shape
    mask = np.full((bsz, all_sz), False)
    if mask_type != "static" or no_overlap or num_mask_ver != 2 or idc_select_ver != 1:
        raise NotImplementedError("only the branch reachable from the shipped configs is restated")
    mask_idcs = []
    rng = None
    for i in range(bsz):
        if seed is not None and epoch is not None and indices is not None:
            seed_i = int(hash((seed, epoch, indices[i].item())) % 1e6)
        else:
            seed_i = None
        rng = np.random.default_rng(seed_i)
        if padding_mask is not None:
            sz = all_sz - padding_mask[i].long().sum().item()
            assert sz >= 0, sz
        else:
            sz = all_sz
        num_mask = int(mask_prob * sz / float(mask_length) + rng.random())
        num_mask = max(min_masks, num_mask)
        lengths = np.full(num_mask, mask_length)
        if sum(lengths) == 0:
            raise ValueError("this should never happens")
        min_len = min(lengths)
        if sz - min_len <= num_mask:
            min_len = sz - num_mask - 1
        mask_idc = rng.choice(sz - min_len, num_mask, replace=False)
        mask_idc = np.asarray([mask_idc[j] + offset for j in range(len(mask_idc)) for offset in range(lengths[j])])
        mask_idc = np.unique(mask_idc[mask_idc < sz])
        if len(mask_idc) >= sz:
            raise ValueError(f"the entire sequence is masked. sz={sz}; mask_idc[mask_idc]; index={indices[i] if indices is not None else None}")
        mask_idcs.append(mask_idc)

    target_len = None
    if require_same_masks:
        target_len = max(len(m) for m in mask_idcs) if add_masks else min(len(m) for m in mask_idcs)

    for i, mask_idc in enumerate(mask_idcs):
        if target_len is not None and len(mask_idc) > target_len:
            mask_idc = rng.choice(mask_idc, target_len, replace=False)
        mask[i, mask_idc] = True
        if target_len is not None and len(mask_idc) < target_len:
            unmasked = np.flatnonzero(~mask[i])
            to_mask = rng.choice(unmasked, target_len - len(mask_idc), replace=False)
            mask[i, to_mask] = True
        if mask_dropout > 0:
            masked = np.flatnonzero(mask[i])
            num_holes = np.rint(len(masked) * mask_dropout).astype(int)
            to_drop = rng.choice(masked, num_holes, replace=False)
            mask[i, to_drop] = False
    return mask


class Mlp(tnn.Module):
    """timm 0.6.12 Mlp."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=tnn.GELU, bias=True, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = tnn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = tnn.Dropout(drop)
        self.fc2 = tnn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = tnn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class DropPath(tnn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        rt = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            rt.div_(keep)
        return x * rt


class FairseqDataclass:
    pass


class FairseqCriterion(tnn.Module):
    def __init__(self, task=None):
        super().__init__()
        self.task = task


class ModelCriterion(FairseqCriterion):
    """fairseq.criterions.model_criterion.ModelCriterion."""

    def __init__(self, task, loss_weights=None, log_keys=None, can_sum=True):
        super().__init__(task)
        self.loss_weights = loss_weights
        self.log_keys = log_keys
        self.can_sum = can_sum

    def forward(self, model, sample, reduce=True):
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
                coef = 1.0 if len(self.loss_weights) == 0 else self.loss_weights[lk]
            except (KeyError, TypeError):
                coef = 1.0
            if coef != 0 and p is not None:
                scaled_losses[lk] = coef * p.float().sum()
        loss = sum(scaled_losses.values())
        if "sample_size" in net_output:
            sample_size = net_output["sample_size"]
        else:
            sample_size = loss.numel()
        if reduce and loss.numel() > 1:
            loss = loss.sum()
        logging_output = {
            "loss": loss.data,
            "ntokens": sample_size,
            "nsentences": sample["id"].numel(),
            "sample_size": sample_size,
            "_world_size": 1,
        }
        for lk in self.log_keys or []:
            if lk in net_output and net_output[lk] is not None:
                if not torch.is_tensor(net_output[lk]) or net_output[lk].numel() == 1:
                    logging_output[lk] = float(net_output[lk])
                elif lk.startswith("_"):
                    logging_output[lk] = net_output[lk]
                else:
                    for i, v in enumerate(net_output[lk]):
                        logging_output[f"{lk}_{i}"] = float(v)
        if len(scaled_losses) > 1:
            for lk, l in scaled_losses.items():
                if l.numel() > 1:
                    l = l.sum()
                logging_output[f"loss_{lk}"] = l.item()
        return loss, sample_size, logging_output


def _populate(module):
    name = module.__name__
    if name == "omegaconf":
        module.II = lambda s: None
        module.MISSING = "???"
    elif name == "fairseq.modules":
        for obj in (TransposeLast, SamePad, Fp32LayerNorm, Fp32GroupNorm, GradMultiply, EMAModule, EMAModuleConfig):
            setattr(module, obj.__name__, obj)
        module.LayerNorm = LayerNorm
    elif name == "fairseq.modules.transformer_sentence_encoder":
        module.init_bert_params = init_bert_params
    elif name == "fairseq.utils":
        module.index_put = index_put
    elif name == "fairseq.models":
        module.BaseFairseqModel = BaseFairseqModel
        module.FairseqEncoder = type("FairseqEncoder", (tnn.Module,), {"__init__": lambda self, d=None: tnn.Module.__init__(self)})
    elif name == "fairseq.dataclass":
        module.FairseqDataclass = FairseqDataclass
    elif name == "fairseq.data.data_utils":
        module.compute_mask_indices = compute_mask_indices
    elif name == "fairseq.criterions.model_criterion":
        module.ModelCriterion = ModelCriterion
        module.ModelCriterionConfig = type("ModelCriterionConfig", (FairseqDataclass,), {})
    elif name == "fairseq.criterions.label_smoothed_cross_entropy":
        module.LabelSmoothedCrossEntropyCriterion = type("LabelSmoothedCrossEntropyCriterion", (FairseqCriterion,), {})
        module.LabelSmoothedCrossEntropyCriterionConfig = type("LabelSmoothedCrossEntropyCriterionConfig",
                                                               (FairseqDataclass,), {})
    elif name == "fairseq.tasks.audio_pretraining":
        module.AudioPretrainingConfig = type("AudioPretrainingConfig", (FairseqDataclass,), {})
        module.AudioPretrainingTask = type("AudioPretrainingTask", (object,), {})
    elif name == "fairseq.models.wav2vec":
        module.Wav2Vec2CtcConfig = type("Wav2Vec2CtcConfig", (FairseqDataclass,), {})
        module.Wav2VecCtc = type("Wav2VecCtc", (BaseFairseqModel,), {})
        module.Wav2VecEncoder = type("Wav2VecEncoder", (tnn.Module,), {})
    elif name == "fairseq.data.audio.raw_audio_dataset":
        module.RawAudioDataset = type("RawAudioDataset", (object,), {})
    elif name == "fairseq.tasks":
        module.FairseqTask = type("FairseqTask", (object,), {})
    elif name == "fairseq.logging.meters":
        module.Meter = type("Meter", (object,), {})
    elif name == "timm.models.vision_transformer":
        module.Mlp = Mlp
        module.DropPath = DropPath
        module.PatchEmbed = type("PatchEmbed", (tnn.Module,), {})
    elif name == "matplotlib":
        module.use = lambda *a, **k: None


_installed = False


def install():
    """Make ``import nn`` (the reference package) work. Idempotent."""
    global _installed
    if _installed:
        return
    _installed = True
    sys.meta_path.insert(0, _StubFinder())

    real_dataclass = dataclasses.dataclass

    def patched_dataclass(cls=None, /, **kw):
        def wrap(c):
            if c.__module__ == "nn" or c.__module__.startswith("nn."):
                kw2 = dict(kw)
                kw2["unsafe_hash"] = True  # Python >= 3.11 rejects unhashable instance defaults
                return real_dataclass(c, **kw2)
            return real_dataclass(c, **kw)

        return wrap if cls is None else wrap(cls)

    dataclasses.dataclass = patched_dataclass
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    install()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nn  # noqa: F401  (the reference package)

    return nn
