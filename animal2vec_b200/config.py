"""Configuration surface of the pretraining path.

The dataclasses keep the reference's field names and defaults so a hydra/YAML config written
for the reference resolves unchanged:
  Data2VecMultiConfig   <- /root/reference/nn/data2vec2.py:56-166
  D2vModalitiesConfig   <- data2vec2.py:50-53
  D2vModalityConfig     <- nn/modalities/base.py:28-72
  D2vAudioConfig        <- nn/modalities/audio.py:29-51
  D2vDecoderConfig      <- nn/modalities/modules.py:34-47
Fields the reference fills through omegaconf ``II("...")`` interpolation (task.sample_rate,
task.conv_feature_layers, common.seed, optimization.max_update, model.num_heads, model.depth)
are plain fields here; :func:`resolve` propagates them the way the interpolation would.
Python >= 3.11 forbids dataclass-instance defaults, hence ``default_factory``.
"""
from __future__ import annotations

import ast
from dataclasses import dataclass, field, fields, is_dataclass
from enum import Enum
from typing import Any, List, Optional, Tuple

SHIPPED_CONV_LAYERS = "[(127, 63, 1)] +[(512, 10, 5)] + [(512, 3, 2)] * 3 + [(512, 3, 1)] + [(512, 2, 1)] * 2"


class Modality(Enum):
    """nn/modalities/modality.py"""

    AUDIO = 1
    IMAGE = 2


@dataclass
class D2vDecoderConfig:
    decoder_dim: int = 384
    decoder_groups: int = 16
    decoder_kernel: int = 5
    decoder_layers: int = 5
    input_dropout: float = 0.1
    add_positions_masked: bool = False
    add_positions_all: bool = False
    decoder_residual: bool = True
    projection_layers: int = 1
    projection_ratio: float = 2.0


@dataclass
class D2vModalityConfig:
    type: Modality = Modality.AUDIO
    prenet_depth: int = 4
    prenet_layerdrop: float = 0
    prenet_dropout: float = 0
    start_drop_path_rate: float = 0
    end_drop_path_rate: float = 0
    num_extra_tokens: int = 0
    init_extra_token_zero: bool = True
    mask_noise_std: float = 0.01
    mask_prob_min: Optional[float] = None
    mask_prob: float = 0.7
    inverse_mask: bool = False
    mask_prob_adjust: float = 0
    keep_masked_pct: float = 0
    mask_length: int = 5
    add_masks: bool = False
    remove_masks: bool = False
    mask_dropout: float = 0.0
    encoder_zero_mask: bool = True
    mask_channel_prob: float = 0.0
    mask_channel_length: int = 64
    ema_local_encoder: bool = False
    local_grad_mult: float = 1.0
    use_alibi_encoder: bool = False
    alibi_scale: float = 1.0
    learned_alibi: bool = False
    alibi_max_pos: Optional[int] = None
    learned_alibi_scale: bool = False
    learned_alibi_scale_per_head: bool = False
    learned_alibi_scale_per_layer: bool = False
    num_alibi_heads: Optional[int] = None  # II("model.num_heads")
    model_depth: Optional[int] = None  # II("model.depth")
    decoder: Optional[D2vDecoderConfig] = field(default_factory=D2vDecoderConfig)


@dataclass
class D2vAudioConfig(D2vModalityConfig):
    type: Modality = Modality.AUDIO
    extractor_mode: str = "layer_norm"
    conv_feature_layers: Optional[str] = None  # II("task.conv_feature_layers")
    sample_rate: Optional[int] = None  # II("task.sample_rate")
    conv_pos_width: int = 95
    conv_pos_groups: int = 16
    conv_pos_depth: int = 5
    conv_pos_pre_ln: bool = False
    sinc_input: bool = True
    apply_window_to_root: bool = False
    sinc_norm: str = "instance"
    use_pswish: bool = False


@dataclass
class D2vModalitiesConfig:
    audio: D2vAudioConfig = field(default_factory=D2vAudioConfig)


@dataclass
class Data2VecMultiConfig:
    loss_beta: float = 0
    loss_scale: Optional[float] = None
    depth: int = 8
    start_drop_path_rate: float = 0
    end_drop_path_rate: float = 0
    num_heads: int = 12
    norm_eps: float = 1e-6
    norm_affine: bool = True
    encoder_dropout: float = 0.1
    post_mlp_drop: float = 0.1
    attention_dropout: float = 0.1
    activation_dropout: float = 0.0
    dropout_input: float = 0.0
    layerdrop: float = 0.0
    embed_dim: int = 768
    mlp_ratio: float = 4
    layer_norm_first: bool = False
    average_top_k_layers: int = 16
    end_of_block_targets: bool = False
    clone_batch: int = 1
    layer_norm_target_layer: bool = False
    batch_norm_target_layer: bool = False
    instance_norm_target_layer: bool = False
    instance_norm_targets: bool = False
    layer_norm_targets: bool = False
    ema_decay: float = 0.999
    ema_same_dtype: bool = True
    log_norms: bool = True
    ema_end_decay: float = 0.9999
    ema_anneal_end_step: int = 300000  # II("optimization.max_update")
    ema_encoder_only: bool = True
    max_update: int = 300000  # II("optimization.max_update")
    modalities: D2vModalitiesConfig = field(default_factory=D2vModalitiesConfig)
    shared_decoder: Optional[D2vDecoderConfig] = None
    min_target_var: float = 0.1
    min_pred_var: float = 0.01
    supported_modality: Optional[Modality] = None
    mae_init: bool = False
    seed: int = 1  # II("common.seed")
    skip_ema: bool = False
    cls_loss: float = 0
    recon_loss: float = 0
    d2v_loss: float = 1
    decoder_group: bool = False
    final_dropout: float = 0.0
    unique_labels: Optional[str] = None
    with_labels: bool = False
    use_focal_loss: bool = False
    sample_rate: int = 8000  # II("task.sample_rate")
    metric_threshold: float = 0.25
    iou_threshold: float = 0.0
    sigma_s: float = 0.1
    maxfilt_s: float = 0.1
    max_duration_s: float = 0.5
    lowP: float = 0.125
    method: str = "avg"
    segmentation_metrics: bool = False
    verbose_tensorboard_logging: bool = False
    conv_feature_layers: str = SHIPPED_CONV_LAYERS  # II("task.conv_feature_layers")
    mixup_prob: float = 0.5
    mixing_window_length: float = 0.1
    source_mixup: float = -1.0
    same_mixup: bool = True
    gain_mode: str = "A_weighting"
    target_mixup: bool = False


def resolve(cfg: Data2VecMultiConfig) -> Data2VecMultiConfig:
    """Fill the fields the reference resolves through omegaconf interpolation."""
    a = cfg.modalities.audio
    if a.conv_feature_layers is None:
        a.conv_feature_layers = cfg.conv_feature_layers
    if a.sample_rate is None:
        a.sample_rate = cfg.sample_rate
    if a.num_alibi_heads is None:
        a.num_alibi_heads = cfg.num_heads
    if a.model_depth is None:
        a.model_depth = cfg.depth
    return cfg


def from_dict(cls, d: Any):
    """Build (nested) config dataclasses from a plain dict such as a loaded YAML ``model:`` node."""
    if d is None or not is_dataclass(cls):
        return d
    if isinstance(d, cls):
        return d
    kw = {}
    names = {f.name: f for f in fields(cls)}
    for k, v in dict(d).items():
        if k.startswith("_"):
            continue
        if k not in names:
            raise KeyError(f"{cls.__name__} has no field {k!r}")
        sub = {"modalities": D2vModalitiesConfig, "audio": D2vAudioConfig, "decoder": D2vDecoderConfig,
               "shared_decoder": D2vDecoderConfig}.get(k)
        kw[k] = from_dict(sub, v) if (sub is not None and isinstance(v, dict)) else v
    return cls(**kw)


def parse_conv_layers(spec: str) -> List[Tuple[int, int, int]]:
    """The reference ``eval``s this string (nn/modalities/audio.py:68); only list/tuple arithmetic is accepted here."""
    tree = ast.parse(spec, mode="eval")
    for node in ast.walk(tree):
        if not isinstance(node, (ast.Expression, ast.BinOp, ast.Add, ast.Mult, ast.List, ast.Tuple, ast.Constant,
                                 ast.Load)):
            raise ValueError(f"conv_feature_layers: unsupported syntax {type(node).__name__}")
    out = eval(compile(tree, "<conv_feature_layers>", "eval"), {"__builtins__": {}})
    for t in out:
        assert len(t) == 3, "invalid conv definition: " + str(t)  # nn/utils.py:1135
    return [tuple(int(v) for v in t) for t in out]


def shipped_large(**kw) -> Data2VecMultiConfig:
    """configs/MeerKAT/a2v_large_pretrain_best.yaml:83-147."""
    audio = D2vAudioConfig(
        sinc_input=True, apply_window_to_root=False, use_pswish=True, sinc_norm="layer_norm", conv_pos_depth=5,
        conv_pos_width=95, conv_pos_groups=16, prenet_depth=8, mask_prob=1.5, mask_length=2, mask_prob_adjust=0.05,
        inverse_mask=False, mask_noise_std=0.01, mask_dropout=0, add_masks=False, ema_local_encoder=False,
        use_alibi_encoder=True, prenet_layerdrop=0, prenet_dropout=0.1, learned_alibi_scale=True,
        learned_alibi_scale_per_head=True,
        decoder=D2vDecoderConfig(input_dropout=0.1, decoder_dim=768, decoder_groups=16, decoder_kernel=7,
                                 decoder_layers=4))
    d = dict(depth=16, embed_dim=1024, num_heads=16, clone_batch=12, ema_decay=0.9997, ema_end_decay=1.0,
             ema_anneal_end_step=300000, ema_encoder_only=False, average_top_k_layers=16,
             instance_norm_target_layer=True, layerdrop=0.0, norm_eps=1e-5, loss_beta=0, loss_scale=None,
             source_mixup=0.5, mixup_prob=1.0, same_mixup=True, mixing_window_length=0.05, gain_mode="A_weighting",
             target_mixup=False, max_update=384230, seed=1, sample_rate=8000,
             modalities=D2vModalitiesConfig(audio=audio))
    d.update(kw)
    return resolve(Data2VecMultiConfig(**d))


def shipped_base(**kw) -> Data2VecMultiConfig:
    """"animal2vec-base" (SURVEY.md): dataclass-default transformer + the shipped recipe for the rest."""
    c = shipped_large(depth=8, embed_dim=768, num_heads=12, average_top_k_layers=8)
    c.modalities.audio.prenet_depth = 4
    c.modalities.audio.num_alibi_heads = None
    c.modalities.audio.model_depth = None
    for k, v in kw.items():
        setattr(c, k, v)
    return resolve(c)


def tiny(**kw) -> Data2VecMultiConfig:
    """Small configuration used by the golden fixtures (tests/golden/make_golden.py): every code path
    of the large recipe, 64-wide feature extractor, 48-wide decoder groups."""
    c = shipped_large(
        embed_dim=128, num_heads=2, depth=2, clone_batch=3, average_top_k_layers=2, ema_anneal_end_step=1000,
        conv_feature_layers="[(127, 63, 1)] +[(64, 10, 5)] + [(64, 3, 2)] * 3 + [(64, 3, 1)] + [(64, 2, 1)] * 2")
    a = c.modalities.audio
    a.prenet_depth, a.conv_pos_depth, a.conv_pos_width, a.conv_pos_groups = 2, 2, 38, 2
    a.conv_feature_layers = None
    a.num_alibi_heads = None
    a.model_depth = None
    a.decoder = D2vDecoderConfig(input_dropout=0.1, decoder_dim=96, decoder_groups=2, decoder_kernel=7,
                                 decoder_layers=2)
    for k, v in kw.items():
        setattr(c, k, v)
    return resolve(c)


@dataclass
class Wav2Vec2CcasFinetuneConfig:
    """nn/wav2vec2.py:39-54 on top of fairseq's Wav2Vec2CtcConfig / Wav2Vec2AsrConfig (third party, absent from
    /root/reference; field names and defaults restated from fairseq @ 920a548, models/wav2vec/wav2vec2_asr.py)."""

    # ---- fairseq Wav2Vec2AsrConfig
    w2v_path: Optional[str] = None
    no_pretrained_weights: bool = False
    dropout_input: float = 0.0
    final_dropout: float = 0.0
    dropout: float = 0.0
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    conv_feature_layers: Optional[str] = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] + [(512,2,2)]"
    encoder_embed_dim: Optional[int] = 768
    apply_mask: bool = False
    mask_length: int = 10
    mask_prob: float = 0.5
    mask_selection: str = "static"
    mask_other: float = 0
    no_mask_overlap: bool = False
    mask_min_space: int = 1
    require_same_masks: bool = True
    mask_dropout: float = 0.0
    mask_channel_length: int = 10
    mask_channel_prob: float = 0.0
    mask_channel_selection: str = "static"
    mask_channel_other: float = 0
    no_mask_channel_overlap: bool = False
    freeze_finetune_updates: int = 0
    feature_grad_mult: float = 0.0
    layerdrop: float = 0.0
    drop_path: float = 0
    mask_channel_min_space: int = 1
    mask_channel_before: bool = False
    normalize: bool = True  # II("task.normalize")
    update_alibi: bool = True
    data: Optional[str] = None  # II("task.data")
    w2v_args: Any = None
    offload_activations: bool = False
    min_params_to_wrap: int = int(1e8)
    checkpoint_activations: bool = False
    ddp_backend: Optional[str] = None  # II("distributed_training.ddp_backend")
    zero_mask: bool = False
    load_ema: bool = False
    layer_decay: float = 1
    # ---- fairseq Wav2Vec2CtcConfig
    blank_weight: float = 0
    blank_mode: str = "add"
    # ---- nn/wav2vec2.py:39-54
    unique_labels: Optional[str] = None  # II("task.unique_labels")
    average_top_k_layers: int = 16
    use_focal_loss: bool = True  # II("criterion.use_focal_loss")
    sample_rate: int = 8000  # II("task.sample_rate")
    mixup_prob: float = 0.5
    mixing_window_length: float = 0.1
    source_mixup: float = -1.0
    same_mixup: bool = True
    target_mixup: bool = True
    gain_mode: str = "A_weighting"
    load_pretrain_weights: bool = False


def shipped_finetune(**kw) -> Wav2Vec2CcasFinetuneConfig:
    """configs/MeerKAT/finetune_mixup_100.yaml:82-120 (model section)."""
    d = dict(freeze_finetune_updates=10000, feature_grad_mult=0.0, apply_mask=True, average_top_k_layers=16,
             mask_prob=0.825, mask_length=4, mask_channel_prob=0.5, mask_channel_length=64, dropout=0.1,
             dropout_input=0.0, activation_dropout=0.1, attention_dropout=0.2, final_dropout=0.0, layerdrop=0.1,
             drop_path=0.0, target_mixup=True, source_mixup=0.5, mixup_prob=1.0, same_mixup=True,
             mixing_window_length=0.05, gain_mode="A_weighting", load_pretrain_weights=False,
             unique_labels="['beep', 'synch', 'sn', 'cc', 'ld', 'oth', 'mo', 'al', 'soc', 'agg', 'eating', 'focal']")
    d.update(kw)
    return Wav2Vec2CcasFinetuneConfig(**d)


def finetune_overrides(model_cfg: Data2VecMultiConfig, ft: Wav2Vec2CcasFinetuneConfig) -> Data2VecMultiConfig:
    """The ``arg_overrides`` Wav2VecEncoderModOut applies to the pretraining config when it rebuilds the model from
    the checkpoint (nn/wav2vec2.py:95-130), plus remove_pretraining_modules' clone_batch = 1 (nn/data2vec2.py:1127)."""
    import copy

    c = copy.deepcopy(model_cfg)
    a = c.modalities.audio
    c.encoder_dropout = c.post_mlp_drop = a.prenet_dropout = ft.dropout
    c.activation_dropout = ft.activation_dropout
    c.dropout_input = ft.dropout_input
    c.attention_dropout = ft.attention_dropout
    c.layerdrop = a.prenet_layerdrop = ft.layerdrop
    c.start_drop_path_rate = c.end_drop_path_rate = ft.drop_path
    a.mask_length, a.mask_prob, a.mask_dropout = ft.mask_length, ft.mask_prob, ft.mask_dropout
    a.mask_channel_length, a.mask_channel_prob = ft.mask_channel_length, ft.mask_channel_prob
    a.encoder_zero_mask = ft.zero_mask
    a.inverse_mask = False
    a.local_grad_mult = ft.feature_grad_mult
    a.learned_alibi_scale = ft.update_alibi
    c.clone_batch = 1
    return resolve(c)


def no_randomness(cfg: Data2VecMultiConfig) -> Data2VecMultiConfig:
    """The deterministic "stage parity" setting: dropouts, mask-token noise and mixup off."""
    cfg.encoder_dropout = cfg.post_mlp_drop = cfg.attention_dropout = cfg.activation_dropout = 0.0
    cfg.dropout_input = 0.0
    a = cfg.modalities.audio
    a.prenet_dropout = 0.0
    a.mask_noise_std = 0.0
    a.decoder.input_dropout = 0.0
    cfg.source_mixup = -1.0
    return cfg
