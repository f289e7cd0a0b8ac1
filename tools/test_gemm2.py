"""Correctness + speed of the CTA-pair GEMM against torch (run with A2V_GEMM_2CTA=1 and =0 to compare)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from animal2vec_b200 import gemm

def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
def bench(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

print("A2V_GEMM_2CTA =", os.environ.get("A2V_GEMM_2CTA", "(default on)"))
g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in [(1000, 256, 64), (515, 512, 256), (42624, 1024, 1024), (48000, 3072, 1024)]:
    a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(n, device="cuda", generator=g)
    r = torch.randn(m, n, device="cuda", generator=g).bfloat16()
    u = a.float() @ w.float().t()
    pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    errs = [rel(gemm.gemm_nt(a, w), u), rel(gemm.gemm_nt(a, w, bias=b), u + b),
            rel(gemm.gemm_nt(a, w, bias=b, act=1), F.gelu(u + b)),
            rel(gemm.gemm_nt(a, w, bias=b, act=1, preact=pre), F.gelu(u + b)), rel(pre, u + b),
            rel(gemm.gemm_nt(a, w, residual=r), u + r.float())]
    print((m, n, k), " ".join(f"{e:.2e}" for e in errs), "OK" if max(errs) < 6e-3 else "FAIL", flush=True)
for (m, n, k) in [(42624, 4096, 1024), (42624, 1024, 4096), (48000, 3072, 1024), (42624, 1024, 1024)]:
    a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(n, device="cuda", generator=g)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    t0 = bench(lambda: gemm.gemm_nt(a, w, out=out))
    t1 = bench(lambda: gemm.gemm_nt(a, w, out=out, bias=b))
    t2 = bench(lambda: gemm.gemm_nt(a, w, out=out, bias=b, act=1, preact=pre))
    t3 = bench(lambda: gemm.gemm_nt(a, w, out=out, residual=pre))
    tt = bench(lambda: torch.matmul(a, w.t(), out=out))
    f = 2 * m * n * k / 1e9
    print((m, n, k), f"plain {f/t0:.0f}  bias {f/t1:.0f}  gelu+preact {f/t2:.0f}  residual {f/t3:.0f}  cuBLAS {f/tt:.0f} TFLOP/s", flush=True)
# TN (weight gradient) products
for (r, m, n) in [(5000, 256, 256), (42624, 1024, 1024), (42624, 4096, 1024), (42624, 1024, 4096), (42624, 3072, 1024)]:
    a = torch.randn(r, m, device="cuda", generator=g).bfloat16()
    b = torch.randn(r, n, device="cuda", generator=g).bfloat16()
    out = torch.ones(m, n, device="cuda")
    gemm.gemm_tn(a, b, out)
    ref = 1 + a.float().t() @ b.float()
    e = rel(out, ref)
    out.zero_()
    t = bench(lambda: gemm.gemm_tn(a, b, out))
    print("tn", (r, m, n), f"rel {e:.2e}", "OK" if e < 1e-5 else "FAIL", f"{2 * r * m * n / t / 1e9:.0f} TFLOP/s", flush=True)
