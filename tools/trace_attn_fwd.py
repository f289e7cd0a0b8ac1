"""Phase timeline of one KV tile of the teacher attention forward (build with EXTRA=-DA2V_ATTN_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops, lib
B, H, L = 24, 16, 2000
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(H)], device="cuda")
scale = torch.ones(H, device="cuda")
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
for _ in range(3):
    ops.attn_fwd(qkv, B, L, H, slopes=slopes, alibi_scale=scale)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.load().a2v_debug_attn_fwd_trace(buf, 64)
t = list(buf)
names = {0: "loop top", 1: "S visible", 2: "scores in registers", 3: "previous PV retired", 4: "P written", 5: "after __syncthreads", 6: "PV issued", 8: "before next-S issue", 9: "next S issued", 10: "next S complete (spin)", 11: "PV block entered", 12: "V landed", 13: "PV issued+committed"}
for k in sorted(names):
    print(f"{t[k] - t[0]:8d}  {names[k]}")
