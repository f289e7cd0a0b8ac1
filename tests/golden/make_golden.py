"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference/nn, imported under oracle/ref_shims.py) on CPU. Only runnable in the build
container (the GPU box has no /root/reference); the resulting .npz files are committed.

    python tests/golden/make_golden.py

Weights come from oracle.a2v_oracle.init_params (seed-deterministic, independent of the
reference's RNG stream) and are loaded into the reference model with strict=True, which also
checks the state-dict key/shape inventory (checkpoint ABI) against the reference.
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

from oracle import a2v_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402


def build_reference_model(cfg: O.OracleConfig, *, dropout: bool, mixup: bool, noise: bool):
    nn = ref_shims.import_reference()
    from nn.data2vec2 import Data2VecMultiConfig, D2vModalitiesConfig, Data2VecMultiModel
    from nn import D2vAudioConfig, D2vDecoderConfig, Modality

    dec = D2vDecoderConfig(decoder_dim=cfg.decoder_dim, decoder_groups=cfg.decoder_groups,
                           decoder_kernel=cfg.decoder_kernel, decoder_layers=cfg.decoder_layers,
                           input_dropout=cfg.decoder_input_dropout if dropout else 0.0)
    audio = D2vAudioConfig(
        type=Modality.AUDIO, prenet_depth=cfg.prenet_depth, prenet_layerdrop=0,
        prenet_dropout=cfg.prenet_dropout if dropout else 0.0,
        mask_noise_std=cfg.mask_noise_std if noise else 0.0, mask_prob=cfg.mask_prob, inverse_mask=False,
        mask_prob_adjust=0.05, mask_length=cfg.mask_length, add_masks=False, mask_dropout=0.0,
        ema_local_encoder=False, use_alibi_encoder=True, learned_alibi_scale=True,
        learned_alibi_scale_per_head=True, num_alibi_heads=cfg.num_heads, model_depth=cfg.depth, decoder=dec,
        extractor_mode="layer_norm", conv_feature_layers=cfg.conv_feature_layers, sample_rate=cfg.sample_rate,
        conv_pos_width=cfg.conv_pos_width, conv_pos_groups=cfg.conv_pos_groups, conv_pos_depth=cfg.conv_pos_depth,
        sinc_input=True, apply_window_to_root=False, sinc_norm="layer_norm", use_pswish=True)
    mcfg = Data2VecMultiConfig(
        loss_beta=0, loss_scale=cfg.loss_scale, depth=cfg.depth, num_heads=cfg.num_heads, norm_eps=cfg.norm_eps,
        encoder_dropout=cfg.encoder_dropout if dropout else 0.0, post_mlp_drop=cfg.post_mlp_drop if dropout else 0.0,
        attention_dropout=cfg.attention_dropout if dropout else 0.0, activation_dropout=0.0, dropout_input=0.0,
        layerdrop=0.0, embed_dim=cfg.embed_dim, mlp_ratio=cfg.mlp_ratio, layer_norm_first=False,
        average_top_k_layers=cfg.average_top_k_layers, clone_batch=cfg.clone_batch,
        instance_norm_target_layer=True, ema_decay=cfg.ema_decay, ema_end_decay=cfg.ema_end_decay,
        ema_anneal_end_step=cfg.ema_anneal_end_step, ema_encoder_only=False, max_update=cfg.ema_anneal_end_step,
        modalities=D2vModalitiesConfig(audio=audio), supported_modality=Modality.AUDIO, seed=cfg.seed,
        unique_labels="['a']", with_labels=False, use_focal_loss=True, sample_rate=cfg.sample_rate,
        conv_feature_layers=cfg.conv_feature_layers, mixup_prob=1.0, mixing_window_length=cfg.mixing_window_length,
        source_mixup=cfg.source_mixup if mixup else -1.0, same_mixup=True, gain_mode="A_weighting",
        target_mixup=False, verbose_tensorboard_logging=False, segmentation_metrics=False)
    model = Data2VecMultiModel.build_model(mcfg, None)
    return model


def load_params(model, params):
    sd = {k: v.clone() for k, v in params.items()}
    sd["_ema"] = {k: v.clone().float() for k, v in params.items() if O.is_teacher_key(k)}
    ref_keys = set(k for k in model.state_dict().keys() if k != "_ema")
    assert ref_keys == set(params.keys()), (sorted(ref_keys - set(params)), sorted(set(params) - ref_keys))
    for k, v in model.state_dict().items():
        if k != "_ema":
            assert tuple(v.shape) == tuple(params[k].shape), (k, v.shape, params[k].shape)
    model.load_state_dict(sd, strict=True)
    # the EMA teacher module itself is refreshed from "_ema" by _load_from_state_dict -> restore()
    tsd = model.ema.model.state_dict()
    assert set(tsd.keys()) == set(sd["_ema"].keys()), "teacher key inventory differs"
    for k, v in tsd.items():
        assert torch.equal(v, sd["_ema"][k]), k


def sub(t, rows=7, cols=5):
    """Deterministic sub-sampling of a (.., R, C) tensor to keep fixtures small."""
    t = t.detach().float()
    if t.dim() == 3:
        return t[:, ::rows, ::cols].contiguous().numpy()
    if t.dim() == 2:
        return t[::rows, ::cols].contiguous().numpy()
    return t.numpy()


def run_case(cfg, *, b, n, num_updates, mixup, seed_w=0, seed_x=0, torch_seed=123):
    model = build_reference_model(cfg, dropout=False, mixup=mixup, noise=False)
    params = O.init_params(cfg, seed_w)
    load_params(model, params)
    model.train()
    model.num_updates = num_updates

    g = torch.Generator().manual_seed(seed_x)
    x = torch.randn(b, n, generator=g)
    x = F.layer_norm(x, (n,))  # task.normalize = true
    ids = torch.arange(b)

    cap = {}
    enc = model.modality_encoders["AUDIO"]
    hooks = []
    for i, layer in enumerate(enc.local_encoder.conv_layers):
        hooks.append(layer.register_forward_hook(lambda m, a, o, i=i: cap.__setitem__(f"fe_layer{i}", o)))
    hooks.append(enc.project_features.register_forward_hook(lambda m, a, o: cap.__setitem__("local_features", o)))
    hooks.append(enc.context_encoder.register_forward_hook(lambda m, a, o: cap.__setitem__("student_prenet_raw", o)))
    hooks.append(model.blocks[-1].register_forward_hook(lambda m, a, o: cap.__setitem__("student_out_raw", o[0])))
    hooks.append(enc.decoder.register_forward_hook(lambda m, a, o: cap.__setitem__("decoder_out", o)))
    orig_mi = enc.make_maskinfo

    def rec_mi(x_, mask_, shape=None):
        mi = orig_mi(x_, mask_, shape)
        cap["mask"] = mi.mask
        cap["ids_keep"] = mi.ids_keep[..., 0]
        return mi

    enc.make_maskinfo = rec_mi
    orig_mt = model.make_targets

    def rec_mt(y, k):
        out = orig_mt(y, k)
        cap["targets"] = out
        return out

    model.make_targets = rec_mt
    if mixup:
        orig_fe = enc.forward

        def rec_fe(features, *a, **k):
            cap["mixed_source"] = features
            return orig_fe(features, *a, **k)

        enc.forward = rec_fe

    torch.manual_seed(torch_seed)
    res = model(source=x, id=ids)
    loss = res["losses"]["AUDIO_regression"].float().sum()
    loss.backward()

    d = cfg.embed_dim
    t = cap["mask"].shape[1]
    gidx = cap["ids_keep"].unsqueeze(-1).expand(-1, -1, d)
    rows = cap["mask"].shape[0]
    out = {
        "num_updates": np.int64(num_updates), "b": np.int64(b), "n": np.int64(n),
        "seed_w": np.int64(seed_w), "seed_x": np.int64(seed_x), "torch_seed": np.int64(torch_seed),
        "mask_packed": np.packbits(cap["mask"].bool().numpy(), axis=1), "T": np.int64(t),
        "loss_sum": np.float64(loss.item()), "sample_size": np.int64(int(res["sample_size"])),
        "pred_var": np.float64(float(res["pred_var"])), "target_var": np.float64(float(res["target_var"])),
        "masked_pct": np.float64(res["masked_pct"]), "ema_decay_x1000": np.float64(res["ema_decay"]),
        "local_features": sub(cap["local_features"]),
        "student_prenet": sub(torch.zeros(rows, t, d).scatter_(1, gidx, cap["student_prenet_raw"].detach())),
        "student_out": sub(torch.zeros(rows, t, d).scatter_(1, gidx, cap["student_out_raw"].detach())),
        "decoder_out": sub(cap["decoder_out"]),
        "targets": sub(cap["targets"]),
    }
    for i in range(len(enc.local_encoder.conv_layers)):
        out[f"fe_layer{i}"] = sub(cap[f"fe_layer{i}"].transpose(1, 2), rows=max(1, cap[f"fe_layer{i}"].shape[2] // 50), cols=3)
    if mixup:
        out["mixed_source"] = cap["mixed_source"].detach().numpy()[:, ::37]
    gn, gh = {}, {}
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        gn[k] = float(p.grad.float().norm())
        gh[k] = p.grad.float().flatten()[:8].numpy()
    keys = sorted(gn.keys())
    out["grad_keys"] = np.array(keys)
    out["grad_norms"] = np.array([gn[k] for k in keys], dtype=np.float64)
    out["grad_heads"] = np.stack([np.pad(gh[k], (0, 8 - len(gh[k]))) for k in keys])

    for h in hooks:
        h.remove()
    # EMA step exactly as the trainer triggers it
    model.set_num_updates(num_updates + 1)
    fp32 = model.ema.fp32_params
    ek = sorted(fp32.keys())
    out["ema_keys"] = np.array(ek)
    out["ema_sums"] = np.array([float(fp32[k].double().sum()) for k in ek])
    out["ema_abs_sums"] = np.array([float(fp32[k].double().abs().sum()) for k in ek])
    out["ema_decay_after"] = np.float64(model.ema.get_decay())
    return out


def masks_case(m, t, seed, updates, ids):
    """Large-recipe masks (M=12, T=2000) through the reference's own clone-id + compute_mask path."""
    cfg = O.tiny_config(clone_batch=m, seed=seed)
    model = build_reference_model(cfg, dropout=False, mixup=False, noise=False)
    enc = model.modality_encoders["AUDIO"]
    from nn import MaskSeed

    out = {}
    for u in updates:
        cap = {}
        orig = enc.make_maskinfo

        def rec(x_, mask_, shape=None):
            cap["mask"] = mask_.clone()
            return orig(x_, mask_, shape)

        enc.make_maskinfo = rec
        with torch.no_grad():
            enc.contextualized_features(torch.zeros(len(ids), t, cfg.embed_dim), None, True, True, clone_batch=m,
                                        mask_seeds=MaskSeed(seed=seed, update=u, ids=torch.tensor(ids)))
        enc.make_maskinfo = orig
        out[f"mask_u{u}"] = np.packbits(cap["mask"].bool().numpy(), axis=1)
    out["meta"] = np.array([m, t, seed, len(ids)], dtype=np.int64)
    out["ids"] = np.array(ids, dtype=np.int64)
    out["updates"] = np.array(updates, dtype=np.int64)
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.tiny_config()
    a = run_case(cfg, b=2, n=16000, num_updates=0, mixup=False)
    np.savez_compressed(os.path.join(HERE, "tiny_u0.npz"), **a)
    print("tiny_u0: loss", a["loss_sum"], "masked_pct", a["masked_pct"], "pred_var", a["pred_var"])
    c = run_case(cfg, b=3, n=16000, num_updates=7, mixup=True, seed_w=3, seed_x=5, torch_seed=99)
    np.savez_compressed(os.path.join(HERE, "tiny_u7_mixup.npz"), **c)
    print("tiny_u7_mixup: loss", c["loss_sum"])
    m = masks_case(12, 2000, 1, [0, 1, 7], [0, 1])
    np.savez_compressed(os.path.join(HERE, "masks_large.npz"), **m)
    hashes = {"hash_1_7_0": int(hash((1, 7, 0)) % 1e6), "clone_hash_1": [int(hash((1, i)) % 1e10) for i in range(3)]}
    print("hash probes", hashes)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
