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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("A2V_BENCH_BATCH", "24")),
                    help="clips per GPU per step (the reference's yaml uses 5 on unnamed GPUs)")
    ap.add_argument("--model", default="large", choices=["large", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    return ap.parse_args()


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
    note = (f"{len(times)} timed step(s) of 1 clip x {cfg.clone_batch} clones, {model} config, fp32, forward+backward+EMA, "
            f"oracle/a2v_oracle.py on torch CPU ({cores} threads), mixup off")
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
    B = args.batch
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
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in tl)
    gemm_flops = sum(f for _, _, f in tl)
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
        "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMM kernels: gemm2cta (CTA pairs, block Linears) + gemm_tcgen05 (all other GEMM / strided-conv launches)",
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     "launches": len(tl), "share_of_step": gemm_ms / ms_total},
        "step_tensor_frac": (flop_clip * value / world / 1e12 / peaks["tflops_sustained"]) if flop_clip else None,
        "loss": log.get("loss"), "gnorm": log.get("gnorm"), "pred_var": log.get("pred_var"),
        "target_var": log.get("target_var"),
        "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                val, cores, done, per, note = time_cpu_port(args.model, 1, 0, 60.0)
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": note}
            except Exception as ex:  # the baseline is a reported number, never the product
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
