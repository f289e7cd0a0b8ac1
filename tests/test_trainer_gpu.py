"""Step driver on the GPU:

* world_size 2: after the bucketed all-reduce every rank holds the SUM of the oracle's per-shard gradients, the packed
  statistics are the global sums, and after the optimizer + EMA step both ranks hold identical parameters. Runs with NCCL
  when the box has two GPUs, otherwise with gloo (CUDA tensors) and both ranks on cuda:0.
* multi-step PretrainTrainer.train_step (two micro-batches per update) against an oracle loop: oracle autograd,
  1 / sum(sample_size), global-norm clip, fairseq Adam with the weight_decay_scale-0 group, cosine warm-up, EMA.
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import a2v_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _clip(n, seed):
    return F.layer_norm(torch.randn(1, n, generator=torch.Generator().manual_seed(seed)), (n,))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_grads(params, x, ids, nu, teacher=None):
    ocfg = O.tiny_config()
    student = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    teacher = O.make_teacher(params) if teacher is None else teacher
    res = O.pretrain_forward(student, teacher, ocfg, x, ids, nu)
    loss = res["losses"]["AUDIO_regression"].sum()
    loss.backward()
    return {k: v.grad for k, v in student.items()}, float(loss), int(res["sample_size"])


def _ddp_worker(rank, world, port, backend, out_dir, precision="fp32"):
    import torch.distributed as dist

    from animal2vec_b200 import config as Cfg
    from animal2vec_b200.engine import PretrainEngine
    from animal2vec_b200.trainer import OptimConfig, PretrainTrainer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        params = O.init_params(O.tiny_config(), 0)
        eng = PretrainEngine(Cfg.no_randomness(Cfg.tiny()), f"cuda:{dev}", precision=precision, init=params)
        tr = PretrainTrainer(eng, OptimConfig(lr=1e-3, warmup_updates=0, max_update=100))
        assert tr.reducer.compress == (precision == "bf16" and backend == "nccl") or backend != "nccl"
        n = 16000
        shards = [(_clip(n, 100 + r), torch.tensor([7 + r])) for r in range(world)]
        x, ids = shards[rank]
        tr.accumulate_and_reduce([(x.cuda(), ids)])
        torch.cuda.synchronize()
        total, loss_sum, ss = None, 0.0, 0
        for xs, ids_s in shards:  # the oracle on EVERY shard, summed
            g, l, s = _oracle_grads(params, xs, ids_s, 0)
            total = g if total is None else {k: total[k] + g[k] for k in g}
            loss_sum, ss = loss_sum + l, ss + s
        worst = max(_rel(eng.S.gview(k), total[k]) for k in total)
        assert worst < (3e-3 if precision == "fp32" else 4e-2), worst
        st = tr.stats.cpu()
        assert abs(float(st[0]) - loss_sum) <= (1e-4 if precision == "fp32" else 1e-2) * loss_sum and int(st[1]) == ss
        # every element of the flat buffer went through a bucket (2 bytes each when the buckets travel as bf16)
        assert tr.reducer.bytes_reduced == eng.S.total * (2 if tr.reducer.compress else 4)
        # one full update, then the replicas must agree bit for bit (same reduced gradient, same optimizer)
        tr.train_step([(x.cuda(), ids)])
        torch.cuda.synchronize()
        mine = torch.stack([eng.S.data.double().sum(), eng.S.data.double().abs().sum(), eng.E.data.double().sum()])
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert all(torch.equal(both[0], b) for b in both), both
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{worst:.3e}")
    finally:
        dist.destroy_process_group()


def test_world2_reduced_gradients_equal_the_sum_of_the_oracles_shard_gradients(tmp_path):
    import torch.multiprocessing as mp

    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    mp.spawn(_ddp_worker, args=(2, _free_port(), backend, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="bf16 gradient buckets travel over NCCL: needs two GPUs")
def test_world2_bf16_compressed_buckets(tmp_path):
    """The production setting: bf16 engine, gradient buckets rounded to bf16 for the all-reduce."""
    import torch.multiprocessing as mp

    mp.spawn(_ddp_worker, args=(2, _free_port(), "nccl", str(tmp_path), "bf16"), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_train_steps_match_an_oracle_optimizer_loop():
    from animal2vec_b200 import config as Cfg
    from animal2vec_b200.engine import PretrainEngine, annealed_decay
    from animal2vec_b200.params import no_decay
    from animal2vec_b200.trainer import OptimConfig, PretrainTrainer, cosine_lr

    ocfg = O.tiny_config()
    cfg = Cfg.no_randomness(Cfg.tiny())
    params = O.init_params(ocfg, 0)
    oc = OptimConfig(lr=2e-3, warmup_updates=2, warmup_init_lr=5e-4, max_update=50, weight_decay=0.01, clip_norm=1.0)
    eng = PretrainEngine(cfg, "cuda", precision="fp32", init=params)
    tr = PretrainTrainer(eng, oc)

    student = {k: v.clone() for k, v in params.items()}
    teacher = O.make_teacher(student)
    m = {k: torch.zeros_like(v) for k, v in student.items()}
    v2 = {k: torch.zeros_like(v) for k, v in student.items()}
    n = 16000
    for step in range(3):
        mbs = [(_clip(n, 10 * step + i), torch.tensor([3 * step + i])) for i in range(2)]
        out = tr.train_step([(x.cuda(), ids) for x, ids in mbs], sync_log=True)
        # ---- oracle: accumulate, scale, clip, Adam, EMA
        grads, loss_sum, ss = None, 0.0, 0
        for x, ids in mbs:
            g, l, s = _oracle_grads(student, x, ids, step, teacher)
            grads = g if grads is None else {k: grads[k] + g[k] for k in g}
            loss_sum, ss = loss_sum + l, ss + s
        grads = {k: g / ss for k, g in grads.items()}
        gnorm = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads.values()))
        coef = min(1.0, oc.clip_norm / (gnorm + 1e-6))
        lr = cosine_lr(oc, step)
        t = step + 1
        bc1, bc2 = 1 - oc.betas[0] ** t, 1 - oc.betas[1] ** t
        for k, p in student.items():
            g = grads[k] * coef
            m[k].mul_(oc.betas[0]).add_(g, alpha=1 - oc.betas[0])
            v2[k].mul_(oc.betas[1]).addcmul_(g, g, value=1 - oc.betas[1])
            if not no_decay(k, tuple(p.shape)):
                p.add_(p, alpha=-oc.weight_decay * lr)
            p.addcdiv_(m[k], v2[k].sqrt().add_(oc.eps), value=-lr * math.sqrt(bc2) / bc1)
        decay = annealed_decay(cfg, t)
        for k in teacher:
            teacher[k].mul_(decay).add_(student[k], alpha=1 - decay)
        # ---- compare
        assert abs(out["loss_sum"] - loss_sum) <= 1e-4 * loss_sum and int(out["sample_size"]) == ss
        assert abs(out["gnorm"] - gnorm) <= 3e-3 * gnorm, (out["gnorm"], gnorm)
        assert abs(out["lr"] - lr) < 1e-12 and abs(out["ema_decay"] - decay * 1000) < 1e-9
        # Adam normalises the step to ~lr per element: compare the UPDATE (p - p0), not p, so the gate is meaningful
        per_key = {k: _rel(eng.S.view(k).cpu() - params[k], student[k] - params[k]) for k in student}
        assert max(per_key.values()) < 5e-2, (step, sorted(per_key.items(), key=lambda kv: -kv[1])[:5])
        upd_g = torch.cat([(eng.S.view(k).cpu() - params[k]).reshape(-1) for k in student])
        upd_o = torch.cat([(student[k] - params[k]).reshape(-1) for k in student])
        assert _rel(upd_g, upd_o) < 1e-2, (step, _rel(upd_g, upd_o))
        ema_g = torch.cat([(eng.E.view(k).cpu() - params[k]).reshape(-1) for k in teacher])
        ema_o = torch.cat([(teacher[k] - params[k]).reshape(-1) for k in teacher])
        # the EMA moves the shadow by (1 - tau) * update ~ 1e-7 of the parameter values: fp32 rounding of
        # tau * E + (1 - tau) * S (different operation order on the two sides) is a visible floor
        p0 = torch.cat([params[k].reshape(-1) for k in teacher])
        err = float((ema_g.double() - ema_o.double()).norm())
        assert err <= 1e-2 * float(ema_o.double().norm()) + 3e-7 * float(p0.double().norm()), (step, err)
