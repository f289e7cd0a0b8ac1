#!/usr/bin/env python
"""Headline benchmark: animal2vec-large data2vec2 pretraining step, 10-s 8 kHz clips per second.

    python bench.py --gpus N --steps K --warmup W            # the B200 implementation (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores

One "step" = one full optimizer update of the hot path on one micro-batch of synthetic clips per GPU:
mixup -> SincNet/conv feature extractor -> 12 multi-mask clones -> student (8+16 ALiBi blocks, decoder)
-> EMA teacher (top-16 instance-normed FFN targets) -> masked regression loss -> backward -> bucketed
gradient all-reduce -> grad-norm clip -> AdamW -> EMA teacher update. Nothing is skipped or cached.
Prints ONE JSON line on rank 0 (contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_CLIP = {"large": 6.12e12, "base": 2.29e12}  # SURVEY.md section 8(d), algorithmic, fwd+bwd+teacher
METRIC = "pretrain 10s-clip samples/sec"
WORKLOAD = ("animal2vec-{model} pretraining step (configs[2]), 10-s 8 kHz clips, M=12 clones, EMA teacher, mixup, dropout, "
            "AdamW + clip + EMA update in the step")
UNIT = "samples/s"
# the other BASELINE.json configs, measured on request (--workload); the driver's default run is the headline above
ALT = {
    "fe": {"metric": "feature extractor fwd+bwd 10s-clip samples/sec",
           "workload": "SincNet + conv feature extractor + project_features forward/backward alone (configs[1]), bf16, "
                       "10-s 8 kHz clips", "flop_per_clip": 3 * 53.5e9, "batch": 64},
    "48k": {"metric": "pretrain 10s-clip samples/sec (48 kHz)",
            "workload": "animal2vec-large pretraining step on 10-s 48 kHz clips (configs[4]): 480 000 samples, T = 12 000 "
                        "frames, ~790 kept tokens per clone, M=12 clones, EMA teacher, mixup, dropout, AdamW + clip + EMA",
            "flop_per_clip": 50.6e12, "batch": 3},
    "finetune": {"metric": "finetune 10s-clip samples/sec",
                 "workload": "animal2vec-large finetuning step (configs[3]): 12-class frame-level multilabel head on the "
                             "mean of the top-16 FFN outputs, focal loss, time + channel masking, layerdrop, source + target "
                             "mixup, full-length (T = 2000) student with gradients, frozen conv extractor, AdamW",
                 "flop_per_clip": 4.93e12, "flop_per_clip_frozen": 1.68e12, "batch": 24},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("A2V_BENCH_BATCH", "0")),
                    help="clips per GPU per step (default: 32 for the headline; the reference's yaml uses 5 on unnamed GPUs)")
    ap.add_argument("--model", default="large", choices=["large", "tiny"])
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "fe", "48k", "finetune"],
                    help="pretrain = the headline (BASELINE configs[2]); fe / 48k / finetune = configs[1] / [4] / [3]")
    ap.add_argument("--frozen", action="store_true", help="finetune workload: the frozen phase (only the head trains)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    return ap.parse_args()


def traffic_per_launch(kernel: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from the committed ncu
    --set full capture of this bench command (profiles/r2_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        for ts, ln in self.lines:
            if not (t0 <= ts <= t1 + 0.2):  # only samples taken while the timed region ran
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU pretraining step
# --------------------------------------------------------------------------------------------------
def time_cpu_port(model: str, steps: int, warmup: int, budget_s: float):
    """Times oracle.a2v_oracle.pretrain_step (forward + backward + EMA, fp32, torch CPU with every host
    thread) on 1 clip (12 clones) of the same configuration. Returns (clips/s, cores, steps_done, note)."""
    import torch
    import torch.nn.functional as F

    from oracle import a2v_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.large_config() if model == "large" else O.tiny_config()
    n = 80000 if model == "large" else 16000
    student = O.init_params(cfg, 0)
    teacher = O.make_teacher(student)
    g = torch.Generator().manual_seed(0)
    x = F.layer_norm(torch.randn(1, n, generator=g), (n,))
    ids = torch.arange(1)
    t_start = time.perf_counter()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.pretrain_step(student, teacher, cfg, x, ids, i, do_mixup=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    per = sum(times) / len(times)
    note = (f"{len(times)} timed step(s) after {warmup} warm-up of 1 clip x {cfg.clone_batch} clones, {model} config, fp32, "
            f"forward+backward+EMA, oracle/a2v_oracle.py (port) on torch CPU ({cores} threads), mixup off. Deviates from "
            f"BASELINE.md section 4 (base config, B=2, the shimmed reference itself): the GPU box has no /root/reference and "
            f"the headline config is large; mean of the timed steps, not best-of-3")
    return 1.0 / per, cores, len(times), per, note


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    val, cores, done, per, note = time_cpu_port(args.model, args.steps, warm, args.cpu_budget_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(model=args.model), "clips_per_step": 1,
                   "sample": "one clip (12 clones) per step: forward + backward + EMA update, the unit SURVEY.md "
                             "section 8(d) config 1 times on the CPU (no optimizer step, mixup off)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": note},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from animal2vec_b200 import config as Cfg
    from animal2vec_b200 import lib as L
    from animal2vec_b200.engine import PretrainEngine
    from animal2vec_b200.trainer import OptimConfig, PretrainTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200: the kernels are sm_100a-only and there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    cfg = Cfg.shipped_large() if args.model == "large" else Cfg.tiny()
    n = 80000 if args.model == "large" else 16000
    B = args.batch or 32  # 99 GB of the 180 GB HBM; +2 % over 24 in a same-box A/B (fixed per-step costs amortised)
    eng = PretrainEngine(cfg, dev, precision="bf16", init_seed=0, rng_seed=1 + rank)
    trainer = PretrainTrainer(eng, OptimConfig())

    # synthetic clips: a different batch every step (pool of `pool` batches), per-clip layer-normed like
    # task.normalize=true; host copies are pinned for the end-to-end leg
    pool = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host = [F.layer_norm(torch.randn(B, n, generator=g), (n,)).pin_memory() for _ in range(pool)]
    devb = [h.to(dev) for h in host]
    ids_of = lambda step: [(step * world + rank) * B + i for i in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        if sampler:
            sampler.mark_end()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    step_no = [0]

    def step_resident(s):
        k = step_no[0]; step_no[0] += 1
        eng.prefetch_mask(trainer.num_updates + 1, ids_of(k + 1), B, n)
        trainer.train_step([(devb[k % pool], ids_of(k))])

    staging = torch.empty(B, n, device=dev)
    loss_host = torch.empty(2, dtype=torch.float64).pin_memory()

    def step_e2e(s):
        k = step_no[0]; step_no[0] += 1
        eng.prefetch_mask(trainer.num_updates + 1, ids_of(k + 1), B, n)
        staging.copy_(host[k % pool], non_blocking=True)          # H2D of this step's clips (pinned)
        out = trainer.train_step([(staging, ids_of(k))])
        loss_host.copy_(out["stats"][0:2], non_blocking=True)      # D2H of loss sum + sample size
        torch.cuda.current_stream().synchronize()

    # Allocator priming (untimed set-up, before the warm-up): the number of kept frames per clone changes from step
    # to step (batch-minimum of the random span masks, 143..152 of 2000), so the caching allocator would keep
    # growing -- cudaMalloc is synchronous -- for several steps. One forward/backward with a mask that keeps MORE
    # frames than any real one (160) makes it hold large-enough blocks from the start; gradients are discarded.
    import numpy as np
    T = eng.frames_for(n)
    prime = np.ones((B * eng.M, T), dtype=bool)
    prime[:, : max(1, int(T * 0.08))] = False
    eng.forward(devb[0], ids_of(0), 0, mask=prime)
    eng.backward()
    eng.zero_grad()
    torch.cuda.synchronize()

    # the clock sampler (an nvidia-smi child process) starts BEFORE the warm-up: its start-up must not
    # land inside the timed region; only samples taken inside the region are reported
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for s in range(args.warmup):
        step_resident(s)
    # ---- device-resident leg (the `value`), with per-GEMM events for the roofline entry
    L.gemm_timeline = []
    launches0 = L.launch_count
    ms_total, clocks = timed(step_resident, args.steps, sampler)
    launches = L.launch_count - launches0
    tl, L.gemm_timeline = L.gemm_timeline, None
    fam = {}
    for a, b, f, kind in tl:
        e = fam.setdefault(kind, [0.0, 0.0, 0])
        e[0] += a.elapsed_time(b); e[1] += f; e[2] += 1
    gemm_ms, gemm_flops, gemm_n = fam.get("linear", [0.0, 0.0, 0])
    log = trainer.log_values()
    # ---- end-to-end leg (host buffers, H2D + D2H inside the timed region)
    step_e2e(0)
    ms_e2e, _ = timed(step_e2e, args.steps)

    peaks = measured_peaks()
    clips = world * B * args.steps
    value = clips / (ms_total / 1e3)
    e2e = clips / (ms_e2e / 1e3)
    flop_clip = FLOP_PER_CLIP.get(args.model)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(model=args.model),
                   "clips_per_gpu_per_step": B, "global_batch": B * world, "samples_per_clip": n,
                   "parallelism": f"dp{world}", "l2_policy": "inputs and activations (>10 GB per step) exceed the 126 MB L2; "
                   "a different synthetic batch every step"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * n * 4, "d2h_bytes_per_step": 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor",
                     "kernel": "gemm2cta_kernel / gemm2cta_tn_kernel (CTA-pair tcgen05 GEMM) + gemm_tcgen05_kernel on the plain "
                               "Linear shapes (QKV, proj, fc1, fc2, decoder / feature projections, strided-conv GEMMs; forward, "
                               "data and weight gradients): every launch with taps = groups = batch = 1",
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": traffic_per_launch("gemm2cta_kernel"),
                     "flops": "algorithmic = executed for these shapes (2 M N K, no padding)",
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     "launches": gemm_n, "share_of_step": gemm_ms / ms_total,
                     "other_tensor_kernels": {
                         k: {"ms_per_step": v[0] / args.steps, "tflops_executed": v[1] / (v[0] / 1e3) / 1e12 if v[0] else None,
                             "launches": v[2], "share_of_step": v[0] / ms_total,
                             "note": ("conv_slab_fwd / conv_slab_wgrad (grouped stride-1 convs); algorithmic FLOPs (the decoder's "
                                      "48-channel groups counted 48 wide; its first layer reads 64-channel groups)")
                             if k == "slab" else "gemm_tcgen05_kernel tap-loop / grouped launches"}
                         for k, v in fam.items() if k != "linear"}},
        "step_tensor_frac": (flop_clip * value / world / 1e12 / peaks["tflops_sustained"]) if flop_clip else None,
        "loss": log.get("loss"), "gnorm": log.get("gnorm"), "pred_var": log.get("pred_var"),
        "target_var": log.get("target_var"),
        "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                val, cores, done, per, note = time_cpu_port(args.model, 3, 1, 90.0)
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": note}
            except Exception as ex:  # the baseline is a reported number, never the product
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (--workload fe | 48k | finetune), one GPU or data parallel
# --------------------------------------------------------------------------------------------------
def run_alt(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from animal2vec_b200 import config as Cfg
    from animal2vec_b200 import lib as L
    from animal2vec_b200 import ops
    from animal2vec_b200.engine import PretrainEngine
    from animal2vec_b200.trainer import OptimConfig, PretrainTrainer, cosine_lr

    spec = ALT[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the kernels are sm_100a-only and there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    B = args.batch or spec["batch"]
    tiny = args.model == "tiny"
    sr = 48000 if args.workload == "48k" else 8000
    n = (96000 if tiny else 480000) if args.workload == "48k" else (16000 if tiny else 80000)
    pool = 3
    g = torch.Generator().manual_seed(4321 + rank)
    host = [F.layer_norm(torch.randn(B, n, generator=g), (n,)).pin_memory() for _ in range(pool)]
    devb = [h.to(dev) for h in host]
    staging = torch.empty(B, n, device=dev)
    out_host = torch.empty(2, dtype=torch.float64).pin_memory()
    extra = {}

    if args.workload == "fe":
        cfg = Cfg.tiny() if tiny else Cfg.shipped_large()
        eng = PretrainEngine(cfg, dev, precision="bf16", init_seed=0)
        eng._refresh_student()
        T = eng.frames_for(n)
        dlf = (torch.randn(B * T, eng.D, device=dev) * 1e-2).bfloat16()
        acc = torch.zeros(2, device=dev, dtype=torch.float64)

        def step(x):
            from types import SimpleNamespace
            c = SimpleNamespace()
            eng.zero_grad()
            lf = eng._fe_forward(x, c, True)
            eng._fe_backward(c, dlf)
            eng._unpack_grads()
            return lf

        def result_of(lf):
            ops.sumsq(eng.S.grad, acc[0:1].zero_())
            return acc
        h2d, d2h = B * n * 4, 16
        extra["hbm_algorithmic_gbs_note"] = "fused lower bound 2.4 MB/clip forward (SURVEY 8d); this chain is unfused"
    elif args.workload == "48k":
        cfg = Cfg.tiny(sample_rate=sr) if tiny else Cfg.shipped_large(sample_rate=sr)
        eng = PretrainEngine(cfg, dev, precision="bf16", init_seed=0, rng_seed=1 + rank)
        trainer = PretrainTrainer(eng, OptimConfig())
        k = [0]

        def step(x):
            i = k[0]; k[0] += 1
            return trainer.train_step([(x, [(i * world + rank) * B + j for j in range(B)])])

        def result_of(out):
            return out["stats"][0:2]
        h2d, d2h = B * n * 4, 16
    else:
        from animal2vec_b200.finetune import FinetuneEngine
        mcfg = Cfg.tiny() if tiny else Cfg.shipped_large()
        ft = Cfg.shipped_finetune(freeze_finetune_updates=(10 ** 9 if args.frozen else 0))
        if tiny:
            ft.average_top_k_layers, ft.mask_channel_length = 2, 16
        fe = FinetuneEngine(mcfg, ft, 12, dev, precision="bf16", init_seed=0, rng_seed=1 + rank, metric_threshold=0.175)
        eng = fe.core
        T = eng.frames_for(n)
        tg = torch.Generator().manual_seed(99 + rank)
        tgt_host = [(torch.rand(B, T, 12, generator=tg) < 0.05).float().pin_memory() for _ in range(pool)]
        tgt_dev = [t.to(dev) for t in tgt_host]
        tgt_stage = torch.empty(B, T, 12, device=dev)
        oc = OptimConfig(lr=3e-5, warmup_updates=2000, warmup_init_lr=1e-10, min_lr=5e-6, max_update=30000, eps=1e-8,
                         weight_decay=0.0, clip_norm=0.0)  # finetune_mixup_100.yaml:62-79
        m1, v1 = torch.zeros_like(eng.S.data), torch.zeros_like(eng.S.data)
        m2, v2 = torch.zeros_like(fe.head), torch.zeros_like(fe.head)
        coef = torch.zeros(2, device=dev, dtype=torch.float32)
        denom = torch.zeros(1, device=dev, dtype=torch.float32)
        sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        k = [0]
        tcur = [None]

        def step(x):
            i = k[0]; k[0] += 1
            fe.zero_grad()
            res = fe.forward(x, tcur[0], training=True)
            fe.backward()
            if world > 1:  # legacy_ddp-style flat all-reduce (finetune yaml), head + encoder
                dist.all_reduce(fe.head_grad)
                if fe.encoder_trainable:
                    dist.all_reduce(eng.S.grad)
            # fairseq: grads * 1 / sample_size (ntokens), adam (no clip in the finetune recipe)
            denom.fill_(float(res["sample_size"] * world))
            sumsq.zero_()
            ops.clip_coef(sumsq, denom, 1.0, 0.0, coef)
            lr = cosine_lr(oc, fe.num_updates)
            fe.num_updates += 1
            kw = dict(lr=lr, beta1=0.9, beta2=0.98, eps=oc.eps, weight_decay=0.0, step=fe.num_updates, grad_scale=coef[0:1])
            if fe.encoder_trainable or not args.frozen:
                ops.adamw_step(eng.S.data, eng.S.grad, m1, v1, eng.S16, **kw)
                eng.mark_student_updated(s16_valid=True)
            ops.adamw_step(fe.head, fe.head_grad, m2, v2, None, **kw)
            return res

        def result_of(res):
            return res["loss_sum"]
        h2d, d2h = B * n * 4 + B * T * 12 * 4, 8
        spec = dict(spec)
        if args.frozen:
            spec["flop_per_clip"] = spec["flop_per_clip_frozen"]
            spec["workload"] += " -- FROZEN phase (freeze_finetune_updates not reached: encoder forward only, head trains)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(steps):
            fn(s_)
        e1.record()
        barrier()
        if sampler:
            sampler.mark_end()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    cnt = [0]

    def resident(s_):
        i = cnt[0]; cnt[0] += 1
        if args.workload == "finetune":
            tcur[0] = tgt_dev[i % pool]
        step(devb[i % pool])

    def e2e(s_):
        i = cnt[0]; cnt[0] += 1
        staging.copy_(host[i % pool], non_blocking=True)
        if args.workload == "finetune":
            tgt_stage.copy_(tgt_host[i % pool], non_blocking=True)
            tcur[0] = tgt_stage
        out = step(staging)
        r = result_of(out)
        out_host[: r.numel()].copy_(r.view(-1)[:2].double(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for s_ in range(args.warmup):
        resident(s_)
    L.gemm_timeline = []
    launches0 = L.launch_count
    ms_total, clocks = timed(resident, args.steps, sampler)
    launches = L.launch_count - launches0
    tl, L.gemm_timeline = L.gemm_timeline, None
    t_ms = sum(a.elapsed_time(b) for a, b, _, _ in tl)
    t_fl = sum(f for _, _, f, _ in tl)
    e2e(0)
    ms_e2e, _ = timed(e2e, args.steps)
    peaks = measured_peaks()
    clips = world * B * args.steps
    value = clips / (ms_total / 1e3)
    line = {
        "metric": spec["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": spec["workload"] + (" [tiny model: smoke run]" if tiny else ""), "clips_per_gpu_per_step": B,
                   "global_batch": B * world, "samples_per_clip": n, "parallelism": f"dp{world}",
                   "l2_policy": "a different synthetic batch every step; activations exceed the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": clips / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "all tcgen05 GEMM / slab-conv launches of the step (executed FLOPs)",
                     "achieved": t_fl / (t_ms / 1e3) / 1e12 if t_ms else None, "peak": peaks["tflops_sustained"],
                     "unit": "TFLOP/s", "frac": (t_fl / (t_ms / 1e3) / 1e12 / peaks["tflops_sustained"]) if t_ms else None,
                     "traffic": None, "share_of_step": t_ms / ms_total if ms_total else None,
                     "peak_source": peaks["source"] + ", sustained figure"},
        "step_tensor_frac": spec["flop_per_clip"] * value / world / 1e12 / peaks["tflops_sustained"] if not tiny else None,
        "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    line.update(extra)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "pretrain":
        run_alt(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
