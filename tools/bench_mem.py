"""Micro-benchmarks of the HBM-bound kernels at the large-config shapes (GB/s of algorithmic bytes)."""
import json
import sys
import os

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


def main():
    out = []
    dev = "cuda"
    B = int(os.environ.get("B", "16"))
    big = B * 12 * 2000
    small = B * 12 * 142
    for name, rows, c, gw, gr, act, two, post, affine in [
        ("posconv LN+GELU", big, 1024, 1024, 1024, 1, False, False, False),
        ("decoder LN+GELU+res", big, 1024, 64, 48, 1, False, True, False),
        ("block LN(x+drop(b))", small, 1024, 1024, 1024, 0, True, False, True),
        ("teacher block LN", B * 2000, 1024, 1024, 1024, 0, True, False, True),
        ("fe LN512+GELU", B * 16000, 512, 512, 512, 1, False, False, True),
        ("fe LN127+PSwish", B * 80000, 128, 128, 127, 2, False, False, True),
    ]:
        a = torch.randn(rows, c, device=dev).bfloat16()
        b = torch.randn(rows, c, device=dev).bfloat16() if two else None
        po = torch.randn(rows, c, device=dev).bfloat16() if post else None
        creal = c // gw * gr
        gamma = torch.randn(creal, device=dev) if affine else None
        beta = torch.randn(creal, device=dev) if affine else None
        al = torch.randn(creal, device=dev) if act == 2 else None
        be = torch.randn(creal, device=dev) if act == 2 else None
        cfg = ops.RowLnCfg(c, 1e-5, act=act, group_width=gw, group_real=gr, drop_b=0.1 if two else 0.0)
        y, m, r = ops.rowln_fwd(cfg, a, b, gamma, beta, al, be, po, seed_b=3)
        nt_f = 2 + (1 if two else 0) + (1 if post else 0)
        ms = timeit(lambda: ops.rowln_fwd(cfg, a, b, gamma, beta, al, be, po, seed_b=3))
        gb = rows * c * 2 * nt_f / 1e9
        out.append({"kernel": "rowln_fwd " + name, "ms": ms, "GBps": gb / ms * 1e3})
        dy = torch.randn(rows, c, device=dev).bfloat16()
        dg = torch.zeros(creal, device=dev) if affine else None
        db_ = torch.zeros(creal, device=dev) if affine else None
        dal = torch.zeros(creal, device=dev) if act == 2 else None
        dbe = torch.zeros(creal, device=dev) if act == 2 else None
        fn = lambda: ops.rowln_bwd(cfg, dy, a, b, gamma, beta, al, be, m, r, seed_b=3, dgamma=dg, dbeta=db_,
                                   dact_alpha=dal, dact_beta=dbe)
        ms = timeit(fn)
        nt_b = 3 + (2 if two else 0)
        gb = rows * c * 2 * nt_b / 1e9
        out.append({"kernel": "rowln_bwd " + name, "ms": ms, "GBps": gb / ms * 1e3})
        del a, b, po, dy, y
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
