import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import gemm
m, n, k = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (54528, 1024, 1024))]
a = torch.randn(m, k, device="cuda").bfloat16()
w = torch.randn(n, k, device="cuda").bfloat16()
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    gemm.gemm_nt(a, w, out=out, block_n=256)
torch.cuda.synchronize()
