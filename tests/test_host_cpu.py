"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/a2v_capi.h declares, the
host-side logic of the drop-in (config surface, mask generation, parameter inventory / flat layout, weight-pack
index maps, LR schedule, registry) and the product's refusal to run without a CUDA device."""
import os
import re

import numpy as np
import pytest
import torch

from animal2vec_b200 import config as Cfg
from animal2vec_b200 import lib, masking
from animal2vec_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    return lib.load()


def test_library_exports_every_declared_symbol(built):
    syms = lib.exported_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(built, s), f"liba2v_sm100.so does not export {s}"
    assert built.a2v_version() >= 1


def test_header_cites_reference_lines():
    txt = open(os.path.join(ROOT, "include", "a2v_capi.h")).read()
    # every entry-point family names the reference call site it replaces
    assert len(re.findall(r"nn/[a-z_/0-9]+\.py:\d+", txt)) >= 15


def test_no_cpu_path():
    from animal2vec_b200 import ops

    with pytest.raises(lib.A2VError):
        ops.colsum(torch.zeros(8, 8), torch.zeros(8))  # CPU tensors are refused, there is no fallback
    if not torch.cuda.is_available():
        from animal2vec_b200.engine import PretrainEngine

        with pytest.raises(Exception):
            PretrainEngine(Cfg.tiny(), "cuda")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "animal2vec_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


# ------------------------------------------------------------------------------------------ masks
def test_masks_bit_exact_against_reference_fixture():
    g = np.load(os.path.join(GOLD, "masks_large.npz"))
    m, t, seed, nb = [int(v) for v in g["meta"]]
    for u in g["updates"]:
        ref = np.unpackbits(g[f"mask_u{int(u)}"], axis=1)[:, :t].astype(bool)
        got = masking.pretrain_mask(seed=seed, update=int(u), ids=g["ids"], batch=nb, frames=t, clone_batch=m,
                                    mask_prob=1.5, mask_length=2)
        assert np.array_equal(got, ref)
        assert len(set(got.sum(1).tolist())) == 1  # require_same_masks
        assert 0.92 < got.mean() < 0.94            # yaml comment: ~93 % masked


@pytest.mark.parametrize("name", ["tiny_u0.npz", "tiny_u7_mixup.npz"])
def test_masks_match_the_step_fixtures(name):
    g = np.load(os.path.join(GOLD, name))
    t = int(g["T"])
    ref = np.unpackbits(g["mask_packed"], axis=1)[:, :t].astype(bool)
    got = masking.pretrain_mask(seed=1, update=int(g["num_updates"]), ids=list(range(int(g["b"]))), batch=int(g["b"]),
                                frames=t, clone_batch=3, mask_prob=1.5, mask_length=2)
    assert np.array_equal(got, ref)


def test_mask_edge_cases():
    # unseeded masks (id=None in the reference's real training runs) still equalise the rows
    m = masking.pretrain_mask(seed=1, update=0, ids=None, batch=2, frames=100, clone_batch=2, mask_prob=0.5, mask_length=5)
    assert m.shape == (4, 100) and len(set(m.sum(1).tolist())) == 1
    # the whole sequence may not be masked
    with pytest.raises(ValueError):
        masking.compute_mask_indices(1, 8, 8.0, 2, seed=1, epoch=0, indices=np.array([0]), min_masks=1)
    # prefetcher returns the same masks as the inline call
    pf = masking.MaskPrefetcher(seed=3, batch=2, frames=400, clone_batch=3, mask_prob=1.5, mask_length=2)
    pf.announce(5, [7, 9])
    a = pf.get(5, [7, 9])
    b = masking.pretrain_mask(seed=3, update=5, ids=[7, 9], batch=2, frames=400, clone_batch=3, mask_prob=1.5,
                              mask_length=2)
    assert np.array_equal(a, b)
    assert np.array_equal(pf.get(6, [7, 9]), masking.pretrain_mask(seed=3, update=6, ids=[7, 9], batch=2, frames=400,
                                                                   clone_batch=3, mask_prob=1.5, mask_length=2))
    pf.close()


def test_hash_probes():
    # SURVEY.md Appendix A (Python 3.12 tuple hashing)
    assert int(hash((1, 7, 0)) % 1e6) == 697984
    assert masking.clone_seed_ids(1, np.array([0]), 4).tolist() == [0, 7385056256, 2121911296, 4514358272]


# ------------------------------------------------------------------------------------------ config / params
def test_config_surface_and_counts():
    c = Cfg.shipped_large()
    assert (c.embed_dim, c.num_heads, c.depth, c.clone_batch, c.average_top_k_layers) == (1024, 16, 16, 12, 16)
    a = c.modalities.audio
    assert (a.prenet_depth, a.conv_pos_depth, a.conv_pos_groups, a.mask_prob, a.mask_length) == (8, 5, 16, 1.5, 2)
    assert a.num_alibi_heads == 16 and a.model_depth == 16 and a.sample_rate == 8000  # II(...) resolution
    assert Cfg.parse_conv_layers(a.conv_feature_layers) == [(127, 63, 1), (512, 10, 5)] + [(512, 3, 2)] * 3 + \
        [(512, 3, 1)] + [(512, 2, 1)] * 2
    with pytest.raises(ValueError):
        Cfg.parse_conv_layers("__import__('os').system('true')")
    shapes = P.student_param_shapes(c)
    total = sum(int(np.prod(s)) for s in shapes.values())
    teacher = sum(int(np.prod(s)) for k, s in shapes.items() if P.is_teacher_key(k))
    assert total == 315_830_026 and teacher == 308_542_480  # SURVEY.md Appendix A [probe]
    nodecay = [k for k, s in shapes.items() if P.no_decay(k, s)]
    assert len(shapes) == 342 and len(nodecay) == 226
    assert sum(int(np.prod(shapes[k])) for k in nodecay) == 340_492
    # dataclass defaults are the reference's (nn/data2vec2.py:56-166)
    d = Cfg.Data2VecMultiConfig()
    assert (d.depth, d.embed_dim, d.num_heads, d.ema_decay, d.ema_end_decay, d.clone_batch) == (8, 768, 12, 0.999, 0.9999, 1)
    c2 = Cfg.from_dict(Cfg.Data2VecMultiConfig, {"depth": 4, "modalities": {"audio": {"prenet_depth": 2,
                       "decoder": {"decoder_dim": 96}}}})
    assert c2.depth == 4 and c2.modalities.audio.prenet_depth == 2 and c2.modalities.audio.decoder.decoder_dim == 96


def test_state_dict_keys_match_reference_fixture():
    g = np.load(os.path.join(GOLD, "tiny_u0.npz"))
    shapes = P.student_param_shapes(Cfg.tiny())
    assert set(str(k) for k in g["grad_keys"]) == set(shapes.keys())
    assert set(str(k) for k in g["ema_keys"]) == set(k for k in shapes if P.is_teacher_key(k))


def test_flat_layout_and_pack_maps():
    cfg = Cfg.tiny()
    shapes, order, shared = P.student_layout(cfg)
    fp = P.FlatParams(shapes, "cpu", with_grad=True, order=order)
    assert order[: len(shared)] == shared and all(o % P.ALIGN == 0 for o in fp.offsets.values())
    lo, hi = fp.range_of(shared)
    assert lo == 0 and hi == sum((int(np.prod(shapes[k])) + 7) // 8 * 8 for k in shared)

    def apply(pk, w):
        out = torch.zeros(int(np.prod(pk.out_shape)))
        flat = w.reshape(-1)
        for idx in np.ndindex(*pk.dims):
            src = pk.in_off + sum(i * s for i, s in zip(idx, pk.in_strides))
            dst = sum(i * s for i, s in zip(idx, pk.out_strides))
            out[dst] = flat[src]
        return out.view(pk.out_shape)

    g_, ng, cg, k = 2, 3, 4, 5
    w = torch.randn(g_ * ng, cg, k)
    f = apply(P.pack_conv_fwd("w", g_, ng, cg, k, ngp=4, cgp=8), w)      # (G*ngp, k*cgp)
    d = apply(P.pack_conv_dgrad("w", g_, ng, cg, k, ngp=4, cgp=8), w)    # (G*cgp, k*ngp)
    for g0 in range(g_):
        for n in range(ng):
            for c in range(cg):
                for j in range(k):
                    assert f[g0 * 4 + n, j * 8 + c] == w[g0 * ng + n, c, j]
                    assert d[g0 * 8 + c, (k - 1 - j) * 4 + n] == w[g0 * ng + n, c, j]
    assert f.abs().sum() == pytest.approx(w.abs().sum().item(), rel=1e-6)  # pads stay zero
    lt = apply(P.pack_linear_t("w", 6, 4), torch.arange(24.0).view(6, 4))
    assert torch.equal(lt, torch.arange(24.0).view(6, 4).t())
    ct = apply(P.pack_col_t("w", 6, cg, k, 8), torch.randn(6, cg, k))
    assert ct.shape == (k * 8, 6)


def test_lr_schedule_and_decay():
    from animal2vec_b200.engine import alibi_slopes, annealed_decay
    from animal2vec_b200.trainer import OptimConfig, cosine_lr

    o = OptimConfig(lr=1e-4, warmup_updates=10, max_update=110)
    assert cosine_lr(o, 0) == 0.0 and cosine_lr(o, 5) == pytest.approx(5e-5) and cosine_lr(o, 10) == pytest.approx(1e-4)
    assert cosine_lr(o, 60) == pytest.approx(5e-5) and cosine_lr(o, 110) == pytest.approx(0.0, abs=1e-12)
    c = Cfg.shipped_large()
    assert annealed_decay(c, 0) == pytest.approx(0.9997) and annealed_decay(c, 300000) == 1.0
    assert annealed_decay(c, 150000) == pytest.approx(0.99985)
    assert np.allclose(alibi_slopes(16), [2 ** (-0.5 * (h + 1)) for h in range(16)])
    s12 = alibi_slopes(12)
    assert np.allclose(s12[:8], [2.0 ** -(i + 1) for i in range(8)])
    assert np.allclose(s12[8:], [2 ** -0.5, 2 ** -1.5, 2 ** -2.5, 2 ** -3.5])


def test_registry_names():
    import animal2vec_b200.criterions  # noqa: F401
    import animal2vec_b200.data2vec2  # noqa: F401
    from animal2vec_b200 import registry

    assert "data2vec_multi" in registry.MODELS and "expanded_model" in registry.CRITERIA
    assert registry.DATACLASSES["data2vec_multi"] is Cfg.Data2VecMultiConfig


def test_unsupported_variants_fail_loudly():
    from animal2vec_b200.engine import PretrainEngine

    c = Cfg.shipped_large()
    c.layer_norm_first = True
    with pytest.raises(NotImplementedError):
        PretrainEngine._check_supported(c)
    c = Cfg.shipped_large()
    c.modalities.audio.sinc_norm = "pcen"
    with pytest.raises(NotImplementedError):
        PretrainEngine._check_supported(c)
    PretrainEngine._check_supported(Cfg.shipped_large())  # the shipped recipe itself is supported


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """Every descriptor struct of include/a2v_capi.h, compiled by gcc, has the same size and field offsets as its ctypes
    mirror (fields compared in declaration order) -- the C-ABI boundary is plain data, so this is the whole contract."""
    import ctypes
    import re
    import subprocess

    from animal2vec_b200 import lib as L
    from animal2vec_b200 import ops

    mirrors = {"a2v_operand": L.Operand, "a2v_gemm_desc": L.GemmDesc, "a2v_conv_desc": L.ConvDesc,
               "a2v_rowln_desc": ops.RowLnDesc, "a2v_attn_desc": ops.AttnDesc, "a2v_relayout_item": ops.RelayoutItem}
    hdr = open(os.path.join(ROOT, "include", "a2v_capi.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "a2v_capi.h"', "int main(void) {"]
    fields = {}
    for name in mirrors:
        end = re.search(r"\}\s*" + name + r"\s*;", hdr_nc)
        assert end, name
        start = hdr_nc.rfind("typedef struct {", 0, end.start())
        body = hdr_nc[start + len("typedef struct {"):end.start()]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                ident = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?\s*$", part.strip())
                assert ident, (name, decl)
                names.append(ident[0])
        fields[name] = names
        prog.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in names:
            prog.append(f'  printf(" %zu", offsetof({name}, {f}));')
        prog.append('  printf("\\n");')
    prog += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        cls = mirrors[name]
        assert ctypes.sizeof(cls) == int(size), (name, ctypes.sizeof(cls), size)
        py_offs = [getattr(cls, f).offset for f, _ in cls._fields_]
        assert py_offs == [int(o) for o in offs], (name, py_offs, offs, fields[name])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first) on the tiny model: one JSON line with the
    contract's keys, the same metric / unit / workload text as the B200 arm, and zero GPU launches."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny",
                          "--steps", "1", "--warmup", "0"], check=True, capture_output=True, text=True, timeout=300).stdout
    line = json.loads(out.strip().splitlines()[-1])
    import bench

    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"]["workload"] == bench.WORKLOAD.format(model="tiny")
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_fairseq_registration_hook_against_a_stub_package(tmp_path):
    """VERDICT r1 item 9 / ADVICE r1: when fairseq is importable the classes derive from its base classes and are
    registered with ITS registrars under the reference's names (nn/data2vec2.py:168, nn/criterions.py:388,
    nn/audio_tasks.py:92, nn/wav2vec2.py:57); registrar errors surface. fairseq cannot be installed here, so a stub
    package with fairseq's registrar contract (name uniqueness + base-class check) stands in, in a subprocess."""
    import subprocess
    import sys
    import textwrap

    pkg = tmp_path / "fairseq"
    for sub, base, reg in (("models", "BaseFairseqModel", "register_model"),
                           ("criterions", "FairseqCriterion", "register_criterion"),
                           ("tasks", "FairseqTask", "register_task")):
        d = pkg / sub
        d.mkdir(parents=True)
        parent = "torch.nn.Module" if sub != "tasks" else "object"
        (d / "__init__.py").write_text(textwrap.dedent(f"""
            import torch
            REGISTRY = {{}}
            class {base}({parent}):
                pass
            def {reg}(name, dataclass=None):
                def deco(cls):
                    if name in REGISTRY:
                        raise ValueError("Cannot register duplicate {sub} ({{}})".format(name))
                    if not issubclass(cls, {base}):
                        raise ValueError("{sub} ({{}}: {{}}) must extend {base}".format(name, cls.__name__))
                    REGISTRY[name] = (cls, dataclass)
                    return cls
                return deco
        """))
    (pkg / "__init__.py").write_text("")
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import fairseq.models as FM, fairseq.criterions as FC, fairseq.tasks as FT
        import animal2vec_b200.data2vec2, animal2vec_b200.criterions, animal2vec_b200.wav2vec2, animal2vec_b200.audio_tasks
        from animal2vec_b200 import registry
        assert set(FM.REGISTRY) == {"data2vec_multi", "wav2vec_ccas_finetune"}, FM.REGISTRY
        assert set(FC.REGISTRY) == {"expanded_model", "finetunecriterion"}, FC.REGISTRY
        assert set(FT.REGISTRY) == {"audio_ccas"} and issubclass(FT.REGISTRY["audio_ccas"][0], FT.FairseqTask), FT.REGISTRY
        assert issubclass(FM.REGISTRY["data2vec_multi"][0], FM.BaseFairseqModel)
        assert issubclass(FC.REGISTRY["expanded_model"][0], FC.FairseqCriterion)
        assert FM.REGISTRY["data2vec_multi"][1] is registry.DATACLASSES["data2vec_multi"]
        assert len(registry.FAIRSEQ_REGISTERED) == 5
        try:
            registry.register_model("data2vec_multi_x")(type("NotAModel", (), {}))   # wrong base class must surface, not be swallowed
        except ValueError as e:
            print("surfaced:", e)
        else:
            raise SystemExit("registrar error was swallowed")
        print("OK")
    """) % (str(tmp_path), ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr


# ------------------------------------------------------------------------------------------ input pipeline (8f-4)
def _synth_clip(seed, n):
    """Same generator as tests/golden/make_golden_labels.py."""
    g = np.random.default_rng(seed)
    wav = (g.standard_normal(n) * 0.1 + 0.02).astype(np.float32)
    k = int(g.integers(3, 9))
    start = np.sort(g.integers(0, n - 4000, k))
    end = start + g.integers(37, 3500, k)
    cat = g.integers(0, 11, k)
    foc = g.integers(0, 2, k)
    return wav, start.astype(np.int64), end.astype(np.int64), cat.astype(np.int64), foc.astype(np.int64)


def _write_dataset(tmp_path, g):
    import wave

    root = tmp_path / "data" / "wav"
    (tmp_path / "data" / "lbl").mkdir(parents=True)
    root.mkdir(parents=True)
    lines = [str(tmp_path / "data")]  # root above the wav/ directory, names relative to it (nn/audio_tasks.py:225-235)
    clips = []
    for i in range(int(g["n_clips"])):
        wav, s, e, c, f = _synth_clip(int(g[f"seed{i}"]), int(g[f"n{i}"]))
        pcm = np.clip(np.round(wav * 32768.0), -32768, 32767).astype("<i2")
        with wave.open(str(root / f"c{i}.wav"), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000); w.writeframes(pcm.tobytes())
        np.savez(tmp_path / "data" / "lbl" / f"c{i}.npz", start_frame_lbl=s, end_frame_lbl=e, lbl_cat=c, foc=f)
        lines.append(f"wav/c{i}.wav\t{len(wav)}")
        clips.append((pcm.astype(np.float32) / 32768.0, s, e, c, f))
    (tmp_path / "train.tsv").write_text("\n".join(lines) + "\n")
    return clips


def test_dataset_frame_targets_match_the_reference_dataset(tmp_path):
    """Manifest + wav + label files -> per-clip layer norm + frame-level multi-hot targets, against the output of the
    reference's own FileAudioLabelDataset.__getitem__ (tests/golden/labels.npz, make_golden_labels.py)."""
    from animal2vec_b200 import audio_tasks as AT

    g = np.load(os.path.join(GOLD, "labels.npz"))
    labels = [str(x) for x in g["labels"]]
    clips = _write_dataset(tmp_path, g)
    cfg = AT.AudioConfigCCAS(data=str(tmp_path), normalize=True, with_labels=True, unique_labels=str(labels),
                             conv_feature_layers=str(g["conv"]), min_sample_size=1000)
    task = AT.AudioTaskCCAS.setup_task(cfg)
    ds = task.load_dataset("train", label_ext="npz")
    assert len(ds) == int(g["n_clips"]) and ds.sizes.tolist() == [int(g[f"n{i}"]) for i in range(len(ds))]
    for i in range(len(ds)):
        item = ds[i]
        shape = tuple(int(v) for v in g[f"target_shape{i}"])
        want = np.unpackbits(g[f"target{i}"], axis=1)[:, : shape[1]].astype(np.int64)
        assert item["target"].shape == shape and np.array_equal(item["target"], want), i
        assert AT.feature_frames(int(g[f"n{i}"]), ds.conv_feature_layers) == shape[0]
        # 16-bit PCM quantisation of the synthetic clip is the only difference from the golden float samples
        assert np.allclose(item["source"][:64].numpy(), g[f"source_head{i}"], atol=2e-3)
        assert abs(float(item["source"].double().sum())) < 1e-2 and \
            abs(float(item["source"].double().pow(2).sum()) - float(g[f"source_sq{i}"])) < 1e-3 * float(g[f"source_sq{i}"])
    batch = ds.collater([ds[0], ds[1]])
    assert batch["net_input"]["source"].shape == (2, 80000) and batch["target"].shape == (2, 2000, 12)
    assert batch["ntokens"] == 4000 and batch["id"].tolist() == [0, 1]
    with pytest.raises(NotImplementedError):
        ds.collater([ds[0], ds[2]])  # unequal clip lengths (random crops) are not on this path
    from animal2vec_b200 import registry
    assert registry.TASKS["audio_ccas"] is AT.AudioTaskCCAS and registry.DATACLASSES["audio_ccas"] is AT.AudioConfigCCAS


def test_reference_checkpoint_dict_to_config_and_inventory():
    """a trainer-format checkpoint dict ({"model": ..., "cfg": {...}}) written by the reference resolves to
    this package's config (field names are the reference's, yaml model node of a2v_large_pretrain_best.yaml) and its
    tensors pass the key / shape inventory, incl. the 4-D alibi_scale of old checkpoints."""
    import dataclasses
    import types

    from animal2vec_b200 import checkpoint as CK
    from oracle import a2v_oracle as O

    base = Cfg.tiny()
    model_node = dataclasses.asdict(base)
    model_node["_name"] = "data2vec_multi"
    model_node["supported_modality"] = "AUDIO"
    model_node["modalities"]["audio"]["type"] = "AUDIO"
    model_node["sample_rate"] = None
    params = O.init_params(O.tiny_config(), 0)
    sd = {k: v.clone() for k, v in params.items()}
    sd[O.ENC + "alibi_scale"] = sd[O.ENC + "alibi_scale"].squeeze(0)  # pre-upgrade layout
    sd["_ema"] = {k: v.clone() for k, v in O.make_teacher(params).items()}
    state = {"model": sd, "cfg": types.SimpleNamespace(model=model_node, task={"sample_rate": 8000, "_name": "audio_ccas"})}
    cfg = CK.model_config_from_checkpoint(state)
    assert cfg.embed_dim == base.embed_dim and cfg.modalities.audio.prenet_depth == base.modalities.audio.prenet_depth
    assert cfg.sample_rate == 8000 and cfg.modalities.audio.sample_rate == 8000
    student, ema = CK.split_model_state(state)
    assert student[O.ENC + "alibi_scale"].dim() == 5 and "_ema" not in student
    CK.check_inventory(cfg, student, ema)
    broken = dict(student)
    broken.pop("blocks.0.mlp.fc1.weight")
    with pytest.raises(KeyError):
        CK.check_inventory(cfg, broken, ema)
    with pytest.raises(ValueError):
        CK.model_config_from_checkpoint({"cfg": {"model": {"_name": "wav2vec2"}}})


def test_constructor_surface_modules_carry_the_reference_parameter_inventory():
    """SURVEY 8b constructors (SincConv, ConvFeatureExtractionModel, AudioEncoder, AltBlock, Decoder1d): same signatures,
    and their parameter names / shapes equal the checkpoint ABI of the modality encoder."""
    import torch.nn as nn

    from animal2vec_b200 import modules as M

    cfg = Cfg.tiny()
    a = cfg.modalities.audio
    mk = lambda dp: M.AltBlock(cfg.embed_dim, cfg.num_heads, cfg.mlp_ratio, qkv_bias=True, drop=cfg.encoder_dropout,
                               attn_drop=cfg.attention_dropout, mlp_drop=cfg.activation_dropout,
                               post_mlp_drop=cfg.post_mlp_drop, drop_path=dp, norm_layer=nn.LayerNorm,
                               layer_norm_first=cfg.layer_norm_first, ffn_targets=True)
    enc = M.AudioEncoder(a, cfg.embed_dim, mk, nn.LayerNorm, cfg.layer_norm_first, {}, None)
    got = {O_ENC + k: v for k, v in M.reference_state_keys(enc).items()}
    shapes = {k: tuple(v) for k, v in P.student_param_shapes(cfg).items() if k.startswith(O_ENC)}
    assert got == shapes, (sorted(set(got) ^ set(shapes))[:10])
    blk = mk(0.0)
    want = {k[len("blocks.0."):]: tuple(v) for k, v in P.student_param_shapes(cfg).items() if k.startswith("blocks.0.")}
    assert M.reference_state_keys(blk) == want
    sc = M.SincConv(127, 63, sample_rate=8000)
    low, band = P.sinc_mel_init(127, 63, 8000)
    assert torch.equal(sc.low_hz_.detach(), low) and torch.equal(sc.band_hz_.detach(), band) and sc.min_band_hz == 127
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 4, cfg.embed_dim))
    with pytest.raises(NotImplementedError):
        M.SincConv(127, 63, learnable_filters=True)


O_ENC = "modality_encoders.AUDIO."
