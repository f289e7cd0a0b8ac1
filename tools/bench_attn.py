"""Micro-benchmark of the attention kernels at the large-config shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import ops


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


B = int(os.environ.get("B", "16"))
H = 16
slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(H)], device="cuda")
scale = torch.ones(H, device="cuda")
for name, batch, L, with_pos, drop in [("teacher", B, 2000, False, 0.0), ("long12000", max(1, B // 8), 12000, False, 0.0), ("student", B * 12, 142, True, 0.1),
                                       ("student148", B * 12, 148, True, 0.1), ("student128", B * 12, 128, True, 0.1)]:
    if os.environ.get("ONLY") and os.environ["ONLY"] not in name:
        continue
    qkv = torch.randn(batch, L, 3 * H * 64, device="cuda").bfloat16()
    pos = None
    if with_pos:
        pos = torch.stack([torch.randperm(2000, device="cuda")[:L].sort().values for _ in range(batch)]).int().contiguous()
    skip = os.environ.get("SKIP_FAR", "0") == "1" and not with_pos
    ms = timeit(lambda: ops.attn_fwd(qkv, batch, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=drop, seed=1,
                                     skip_far_keys=skip))
    fl = 4.0 * L * L * 64 * H * batch
    print(json.dumps({"kernel": "attn_fwd " + name, "ms": ms, "tflops": fl / ms / 1e9}))
    if L <= 160:
        out, lse = ops.attn_fwd(qkv, batch, L, H, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=drop, seed=1)
        dout = torch.randn_like(out)
        dsc = torch.zeros(H, device="cuda")
        ms = timeit(lambda: ops.attn_bwd(dout, qkv, out, lse, batch, L, H, pos=pos, slopes=slopes, alibi_scale=scale,
                                         dalibi_scale=dsc, drop_p=drop, seed=1))
        print(json.dumps({"kernel": "attn_bwd " + name, "ms": ms, "tflops": 2.5 * fl / ms / 1e9}))
