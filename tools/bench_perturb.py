"""Which part of bench.py's instrumentation perturbs the timed region? Same resident step, four conditions."""
import os, sys, time, threading, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from animal2vec_b200 import config as Cfg, lib as L
from animal2vec_b200.engine import PretrainEngine
from animal2vec_b200.trainer import PretrainTrainer, OptimConfig
import bench

B = int(os.environ.get("B", "24")); n = 80000
eng = PretrainEngine(Cfg.shipped_large(), "cuda", precision="bf16")
tr = PretrainTrainer(eng, OptimConfig())
xs = [F.layer_norm(torch.randn(B, n), (n,)).cuda() for _ in range(4)]
k = [0]
def step():
    i = k[0]; k[0] += 1
    eng.prefetch_mask(tr.num_updates + 1, list(range((i + 1) * B, (i + 2) * B)), B, n)
    tr.train_step([(xs[i % 4], list(range(i * B, (i + 1) * B)))])
def timed(steps=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): step()
    t_enq = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, t_enq * 1e3 / steps
for _ in range(3): step()
print("plain            ms/step %.1f (host enqueue %.1f)" % timed())
L.gemm_timeline = []
print("gemm_timeline    ms/step %.1f (host enqueue %.1f)" % timed())
L.gemm_timeline = None
s = bench.ClockSampler(0); s.start(); time.sleep(1.0)
print("nvidia-smi -lms  ms/step %.1f (host enqueue %.1f)" % timed())
print(s.stop())
print("plain again      ms/step %.1f (host enqueue %.1f)" % timed())
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = [False]; samples = []
def poll():
    while not stop[0]:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.1)
th = threading.Thread(target=poll, daemon=True); th.start()
print("nvml thread      ms/step %.1f (host enqueue %.1f)" % timed())
stop[0] = True; th.join()
print(len(samples), samples[:3])
print("plain 12 steps   ms/step %.1f (host enqueue %.1f)" % timed(12))
