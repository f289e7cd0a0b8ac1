"""Report every synchronising torch call inside one pretraining step (torch.cuda.set_sync_debug_mode)."""
import os, sys, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from animal2vec_b200 import config as Cfg
from animal2vec_b200.engine import PretrainEngine
from animal2vec_b200.trainer import PretrainTrainer

B = int(os.environ.get("B", "8"))
eng = PretrainEngine(Cfg.shipped_large(), "cuda", precision="bf16")
tr = PretrainTrainer(eng)
x = F.layer_norm(torch.randn(B, 80000), (80000,)).cuda()
ids = lambda k: [k * B + i for i in range(B)]
for i in range(2):
    tr.train_step([(x, ids(i))])
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
warnings.simplefilter("always")
import traceback
_orig = warnings.showwarning
def show(message, category, filename, lineno, file=None, line=None):
    print("SYNC:", message)
    for l in traceback.format_stack()[-9:-2]:
        print("   ", l.strip().replace("\n", " | ")[:200])
warnings.showwarning = show
eng.prefetch_mask(tr.num_updates, ids(2), B, 80000)
time.sleep(1.0)
tr.train_step([(x, ids(2))])
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
print("done")
