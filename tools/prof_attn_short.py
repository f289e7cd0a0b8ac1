"""One forward (+ backward) of the student attention at the large-config shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops
B, H, L = 24 * 12, 16, int(os.environ.get("L", "148"))
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(H)], device="cuda")
scale = torch.ones(H, device="cuda")
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
pos = torch.stack([torch.randperm(2000, device="cuda")[:L].sort().values for _ in range(B)]).int().contiguous()
for _ in range(3):
    out, lse = ops.attn_fwd(qkv, B, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=1)
    dqkv = ops.attn_bwd(torch.randn_like(out), qkv, out, lse, B, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=1)
torch.cuda.synchronize()
