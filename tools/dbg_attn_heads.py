"""Per-head error of the windowed (stream) and full (flash) attention forward against fp32 torch math."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops
torch.manual_seed(0)
batch, seq, heads = 2, 2000, 16
d = heads * 64
qkv = torch.randn(batch, seq, 3 * d, device="cuda").bfloat16()
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(heads)], device="cuda")
scale = torch.ones(heads, device="cuda")
q, k, v = qkv.float().view(batch, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
s = (q * 64 ** -0.5) @ k.transpose(-1, -2)
pos = torch.arange(seq, device="cuda").float()
s = s - slopes.view(1, heads, 1, 1) * (pos[:, None] - pos[None, :]).abs()
ref = (s.softmax(-1) @ v)  # b h l d
full, _ = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale)
fast, _ = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale, skip_far_keys=True)
for name, o in (("full", full), ("fast", fast)):
    o = o.float().view(batch, seq, heads, 64).permute(0, 2, 1, 3)
    e = ((o - ref).pow(2).sum((0, 2, 3)) / ref.pow(2).sum((0, 2, 3))).sqrt()
    print(name, " ".join(f"{x:.2e}" for x in e.tolist()))
