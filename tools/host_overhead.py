"""Host enqueue time vs device time of one pretraining step (is the step launch-bound anywhere?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from animal2vec_b200 import config as Cfg, lib as L
from animal2vec_b200.engine import PretrainEngine
from animal2vec_b200.trainer import PretrainTrainer

B = int(os.environ.get("B", "24"))
eng = PretrainEngine(Cfg.shipped_large(), "cuda", precision="bf16")
tr = PretrainTrainer(eng)
x = F.layer_norm(torch.randn(B, 80000), (80000,)).cuda()
ids = lambda k: [k * B + i for i in range(B)]
for i in range(3):
    eng.prefetch_mask(tr.num_updates + 1, ids(i + 1), B, 80000)
    tr.train_step([(x, ids(i))])
torch.cuda.synchronize()
for i in range(3, 7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count
    t0 = time.perf_counter()
    e0.record()
    eng.prefetch_mask(tr.num_updates + 1, ids(i + 1), B, 80000)
    tr.train_step([(x, ids(i))])
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3 * (t1 - t0):7.1f} ms, device {e0.elapsed_time(e1):7.1f} ms, wall {1e3 * (t2 - t0):7.1f} ms, "
          f"{L.launch_count - n0} launches -> {1e6 * (t1 - t0) / (L.launch_count - n0):.1f} us/launch")
# phase split of the host time
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
tr.train_step([(x, ids(8))])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
