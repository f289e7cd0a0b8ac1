"""The drop-in Python surface on the GPU: Data2VecMultiModel.build_model ->
ExpandedModelCriterion(model, sample) -> loss.backward() -> optimizer -> set_num_updates -> state_dict round trip,
against the reference's golden vectors (tests/golden/tiny_u0.npz) and the PretrainEngine / PretrainTrainer path."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import a2v_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _model(precision="fp32", seed_w=0):
    from animal2vec_b200 import config as Cfg
    from animal2vec_b200.data2vec2 import Data2VecMultiModel

    os.environ["A2V_PRECISION"] = precision
    try:
        model = Data2VecMultiModel.build_model(Cfg.no_randomness(Cfg.tiny()), task=None)
    finally:
        del os.environ["A2V_PRECISION"]
    params = O.init_params(O.tiny_config(), seed_w)
    state = {k: v.clone() for k, v in params.items()}
    state["_ema"] = {k: v.clone() for k, v in O.make_teacher(params).items()}
    missing = model.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model, params


def _sample(g):
    b, n = int(g["b"]), int(g["n"])
    x = F.layer_norm(torch.randn(b, n, generator=torch.Generator().manual_seed(int(g["seed_x"]))), (n,))
    ids = torch.arange(b)
    return {"id": ids, "net_input": {"source": x.cuda(), "id": ids}}


def test_model_criterion_backward_state_dict_against_reference_golden():
    from animal2vec_b200.criterions import ExpandedModelCriterion

    g = dict(np.load(os.path.join(GOLD, "tiny_u0.npz"), allow_pickle=False))
    model, params = _model("fp32", int(g["seed_w"]))
    model.train()
    crit = ExpandedModelCriterion(task=None, log_keys=["ema_decay", "target_var", "pred_var"])
    sample = _sample(g)
    torch.manual_seed(int(g["torch_seed"]))
    loss, sample_size, log = crit(model, sample)
    assert int(sample_size) == int(g["sample_size"])
    assert abs(float(loss) - float(g["loss_sum"])) / float(g["loss_sum"]) < 1e-3
    assert abs(log["pred_var"] - float(g["pred_var"])) / float(g["pred_var"]) < 1e-3
    assert abs(log["target_var"] - float(g["target_var"])) / float(g["target_var"]) < 1e-3
    loss.backward()
    named = dict(model.named_parameters())
    keys = [str(k) for k in g["grad_keys"]]
    assert set(keys) == set(named)  # parameter names == the reference's
    for k, nrm in zip(keys, g["grad_norms"]):
        gr = named[k].grad
        assert gr is not None and abs(gr.norm().item() - nrm) <= 3e-3 * nrm + 1e-7, (k, gr.norm().item(), nrm)
    # optimizer parameter tagging (nn/data2vec2.py:318-322)
    tagged = [k for k, p in named.items() if getattr(p, "optim_overrides", None)]
    assert all(len(named[k].shape) == 1 or k.endswith(".bias") or "alibi_scale" in k or "p_swish" in k for k in tagged)
    assert len(tagged) > 0

    # EMA step through the trainer hook, then the checkpoint ABI
    model.set_num_updates(1)
    sd = model.state_dict()
    ek = [str(k) for k in g["ema_keys"]]
    assert set(sd["_ema"].keys()) == set(ek)
    assert set(k for k in sd if k != "_ema") == set(keys)
    for k, s, sa in zip(ek, g["ema_sums"], g["ema_abs_sums"]):
        assert abs(float(sd["_ema"][k].double().sum()) - s) <= 2e-6 * sa + 1e-9, k
    # round trip into a fresh model: identical forward
    model2, _ = _model("fp32", 5)  # different weights first
    model2.load_state_dict(sd)
    model2.train()
    model2.num_updates = model.num_updates = 3
    torch.manual_seed(1)
    l1, _, _ = crit(model, sample)
    torch.manual_seed(1)
    l2, _, _ = crit(model2, sample)
    assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1))
    # 4-D alibi_scale of an old checkpoint is upgraded (nn/modalities/base.py:152-157); trainer file layout accepted
    sd_old = {k: v for k, v in sd.items()}
    sd_old[O.ENC + "alibi_scale"] = sd[O.ENC + "alibi_scale"].squeeze(0)
    sd_old["_ema"] = dict(sd["_ema"])
    sd_old["_ema"][O.ENC + "alibi_scale"] = sd["_ema"][O.ENC + "alibi_scale"].squeeze(0)
    model2.load_state_dict({"model": sd_old, "cfg": None})
    torch.manual_seed(1)
    l3, _, _ = crit(model2, sample)
    assert abs(float(l1) - float(l3)) <= 1e-6 * abs(float(l1))


def test_two_optimizer_steps_with_outer_zero_grad_match_the_engine_path():
    """an outer optimizer's zero_grad(set_to_none=True) must not let gradients pile up in the
    engine's flat buffer. Two SGD steps through criterion(model, sample) + optimizer.zero_grad() equal two steps of
    the bare PretrainEngine with explicit zero_grad()."""
    from animal2vec_b200 import config as Cfg
    from animal2vec_b200.criterions import ExpandedModelCriterion
    from animal2vec_b200.engine import PretrainEngine

    g = dict(np.load(os.path.join(GOLD, "tiny_u0.npz"), allow_pickle=False))
    model, params = _model("fp32", 0)
    model.train()
    crit = ExpandedModelCriterion(task=None)
    sample = _sample(g)
    opt = torch.optim.SGD(model.parameters(), lr=1e-6)
    eng = PretrainEngine(Cfg.no_randomness(Cfg.tiny()), "cuda", precision="fp32", init=params)
    x, ids = sample["net_input"]["source"], sample["id"]
    for step in range(2):
        opt.zero_grad()  # set_to_none=True: the views are dropped
        assert all(p.grad is None for p in model.parameters())
        loss, _, _ = crit(model, sample)
        loss.backward()
        eng.zero_grad()
        res = eng.forward(x, ids, step)
        eng.backward()
        assert abs(float(loss) - float(res["loss_sum"])) <= 1e-5 * abs(float(loss)), step
        gm = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        ge = torch.cat([eng.S.gview(k).reshape(-1) for k, _ in model.named_parameters()])
        assert _rel(gm, ge) < 1e-4, (step, _rel(gm, ge))
        opt.step()
        model.set_num_updates(step + 1)
        eng.S.data.add_(eng.S.grad, alpha=-1e-6)
        eng.mark_student_updated()
        eng.ema_step(step + 1)
    assert _rel(model.engine.S.data, eng.S.data) < 1e-6
    assert _rel(model.engine.E.data, eng.E.data) < 1e-6
    # model.zero_grad() clears the flat buffer and keeps the views attached
    model.zero_grad()
    assert float(model.engine.S.grad.abs().sum()) == 0.0
    assert all(p.grad is not None and float(p.grad.abs().sum()) == 0.0 for p in model.parameters())


def test_dtype_or_device_moves_are_refused():
    model, _ = _model("bf16", 0)
    for bad in (lambda m: m.half(), lambda m: m.bfloat16(), lambda m: m.cpu(), lambda m: m.to(torch.float64)):
        with pytest.raises(RuntimeError):
            bad(model)
    assert model.cuda() is model and model.float() is model  # same-device / same-dtype moves are no-ops
    p = next(model.parameters())
    assert p.data_ptr() == model.engine.S.view(next(iter(model.engine.S.shapes))).data_ptr()


def test_extract_features_and_remove_pretraining_modules_match_reference_fixture():
    """README inference contract (README.md:69-121) / finetune hand-over (nn/data2vec2.py:1112-1142): eval-mode
    extract_features of the student against the reference's own output (tests/golden/tiny_features.npz)."""
    g = dict(np.load(os.path.join(GOLD, "tiny_features.npz"), allow_pickle=False))
    model, params = _model("fp32", 0)
    model.remove_pretraining_modules(modality="AUDIO")
    sd = model.state_dict()
    assert "_ema" not in sd and not any(".decoder." in k for k in sd)
    model.eval()
    b, n = int(g["b"]), int(g["n"])
    x = F.layer_norm(torch.randn(b, n, generator=torch.Generator().manual_seed(int(g["seed_x"]))), (n,))
    with torch.no_grad():
        res = model.extract_features(x.cuda(), mode="AUDIO", mask=False)
    assert set(res) == {"x", "linear_eval_projection", "padding_mask", "layer_results", "mask"}
    assert len(res["layer_results"]) == int(g["n_layers"])

    def sub(t):
        return t.detach().float().cpu()[:, ::7, ::5]

    assert _rel(sub(res["x"]), g["x"]) < 1e-3, _rel(sub(res["x"]), g["x"])
    for i in range(int(g["n_layers"])):
        assert _rel(sub(res["layer_results"][i]), g[f"layer{i}"]) < 1e-3, i
    with pytest.raises(RuntimeError):
        model.train()
        model(x.cuda(), id=torch.arange(b))  # pretraining forward is gone


def test_load_pretrained_from_a_reference_format_checkpoint():
    """{"model": state incl. "_ema" (4-D alibi_scale), "cfg": {"model": yaml node, "task": ...}} -> model;
    same loss as an engine built from the same tensors; a diverged teacher in "_ema" is honoured."""
    import dataclasses

    from animal2vec_b200 import checkpoint as CK
    from animal2vec_b200 import config as Cfg
    from animal2vec_b200.engine import PretrainEngine

    g = dict(np.load(os.path.join(GOLD, "tiny_u0.npz"), allow_pickle=False))
    params = O.init_params(O.tiny_config(), 0)
    node = dataclasses.asdict(Cfg.no_randomness(Cfg.tiny()))
    node.update({"_name": "data2vec_multi", "supported_modality": "AUDIO"})
    node["modalities"]["audio"]["type"] = "AUDIO"
    sd = {k: v.clone() for k, v in params.items()}
    sd[O.ENC + "alibi_scale"] = sd[O.ENC + "alibi_scale"].squeeze(0)
    teacher = {k: v * 1.01 for k, v in O.make_teacher(params).items()}  # a teacher that has drifted from the student
    sd["_ema"] = {k: v.clone() for k, v in teacher.items()}
    model = CK.load_pretrained({"model": sd, "cfg": {"model": node, "task": {"sample_rate": 8000}}}, precision="fp32")
    model.train()
    sample = _sample(g)
    res = model(**sample["net_input"])
    eng = PretrainEngine(Cfg.no_randomness(Cfg.tiny()), "cuda", precision="fp32", init=params)
    eng.load_teacher(teacher)
    ref = eng.forward(sample["net_input"]["source"], sample["id"], 0, need_grad=False)
    a, b = float(res["losses"]["AUDIO_regression"]), float(ref["loss_sum"])
    assert abs(a - b) <= 1e-6 * abs(b)
    assert abs(a - float(g["loss_sum"])) > 1e-4 * abs(a)  # the drifted teacher changes the targets
