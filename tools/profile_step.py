"""In-situ per-kernel timing of one pretraining step (CUDA events around every C-ABI launch, warm caches,
real stream order) -- complements the cold-cache, serialised ncu launch list."""
import collections, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from animal2vec_b200 import config as Cfg, lib as L
from animal2vec_b200.engine import PretrainEngine
from animal2vec_b200.trainer import PretrainTrainer

B = int(os.environ.get("B", "16"))
eng = PretrainEngine(Cfg.shipped_large(), "cuda", precision="bf16")
tr = PretrainTrainer(eng)
x = F.layer_norm(torch.randn(B, 80000), (80000,)).cuda()
for i in range(2):
    tr.train_step([(x, list(range(B)))])
torch.cuda.synchronize()
L.op_timeline = []
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_step([(x, list(range(B)))])
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
tl, L.op_timeline = L.op_timeline, None
agg = collections.defaultdict(lambda: [0, 0.0])
for name, a, b in tl:
    agg[name][0] += 1
    agg[name][1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
print(f"step {e0.elapsed_time(e1):.1f} ms (wall {wall:.1f}); sum of kernel spans {tot:.1f} ms; {len(tl)} launches")
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for k, v in rows[:int(os.environ.get("TOP", "45"))]:
    print(f"{v[1]:8.2f} ms {100 * v[1] / tot:5.1f}% n={v[0]:4d} avg {v[1] / v[0]:7.3f}  {k}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({k: v for k, v in rows}, open("gpurun_out/profile_step.json", "w"), indent=1)
