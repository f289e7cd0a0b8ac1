"""Phase timeline (SM clocks) of four consecutive work items of CTA 0 of the short-sequence attention forward.
Build with EXTRA=-DA2V_ATTN_TRACE (make -C animal2vec_b200/csrc clean all EXTRA=-DA2V_ATTN_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops, lib
B, H, L = 288, 16, int(os.environ.get("L", "148"))
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(H)], device="cuda")
scale = torch.ones(H, device="cuda")
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
pos = torch.stack([torch.randperm(2000, device="cuda")[:L].sort().values for _ in range(B)]).int().contiguous()
for _ in range(3):
    ops.attn_fwd(qkv, B, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=1)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
lib.load().a2v_debug_attn_short_trace(buf, 128)
t = list(buf)
names = {0: "row: item start", 1: "row: after bar.sync (kp written)", 2: "row: S visible", 3: "row: pass 1 done", 4: "row: pass 2 done (P written)",
         5: "row: O visible", 6: "row: epilogue done", 8: "ctrl: S retired", 9: "ctrl: P arrived", 10: "ctrl: PV issued",
         11: "ctrl: PV retired", 12: "ctrl: next S issued"}
t0 = t[0]
for it in range(4):
    print(f"--- item {6 + it} of CTA 0 ({'full tile' if (6 + it) % 2 == 0 else 'remainder tile'})")
    ev = sorted((t[it * 16 + k] - t0, names[k]) for k in names if t[it * 16 + k])
    for c, n in ev:
        print(f"{c:8d}  {n}")
