"""Phase timeline (SM clock cycles) of one head of the student attention backward; needs a build with
EXTRA=-DA2V_ATTN_TRACE (make -C animal2vec_b200/csrc EXTRA=-DA2V_ATTN_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops, lib

B, H, L = 24 * 12, 16, 148
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(H)], device="cuda")
scale = torch.ones(H, device="cuda")
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
pos = torch.stack([torch.randperm(2000, device="cuda")[:L].sort().values for _ in range(B)]).int().contiguous()
out, lse = ops.attn_fwd(qkv, B, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=1)
dout = torch.randn_like(out)
dsc = torch.zeros(H, device="cuda")
for _ in range(3):
    ops.attn_bwd(dout, qkv, out, lse, B, L, H, pos=pos, slopes=slopes, alibi_scale=scale, dalibi_scale=dsc, drop_p=0.1, seed=1)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
h = lib.load()
h.a2v_debug_attn_trace(buf, 64)
t = list(buf)
names = {0: "C loop top", 2: "C loads landed", 3: "C S0 issued", 4: "C P0 ready", 5: "C dP0/dV0 issued",
         6: "C dS0 ready", 7: "C S1/dQ0/dK0 issued", 9: "C P1 ready", 10: "C dP1/dV1 issued", 11: "C dS1 ready",
         12: "C dQ1/dK1 issued", 13: "C head retired, next loads issued", 16: "R loop top", 17: "R delta done",
         18: "R tile0 start", 19: "R S0 visible", 20: "R P0 written",
         21: "R dP0 visible", 22: "R dS0 written", 23: "R dQ0 visible", 24: "R dQ0 stored", 26: "R tile1 start", 27: "R S1 visible",
         28: "R P1 written", 29: "R dP1 visible", 30: "R dS1 written", 31: "R dQ1 visible", 32: "R dQ1 stored",
         39: "R dK/dV final", 40: "R dK/dV stored"}
t0 = min(v for k, v in enumerate(t) if k in names and v)
for k, v in sorted(((k, t[k]) for k in names if t[k]), key=lambda kv: kv[1]):
    print(f"{v - t0:8d}  {names[k]}")
