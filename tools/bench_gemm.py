"""Micro-benchmark of the tcgen05 GEMM (CUDA events, L2-exceeding operand rotation)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal2vec_b200 import gemm


def bench(fn, iters=20, warm=3):
    for _ in range(warm):
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
    res = []
    for (m, n, k, bn) in [(54528, 1024, 1024, 128), (54528, 1024, 1024, 256), (54528, 4096, 1024, 256), (54528, 1024, 4096, 256),
                          (64000, 3072, 1024, 256), (8192, 8192, 8192, 256), (8192, 8192, 8192, 128)]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = torch.randn(n, k, device="cuda").bfloat16()
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        ms = bench(lambda: gemm.gemm_nt(a, w, out=out, block_n=bn))
        ms_t = bench(lambda: torch.matmul(a, w.t(), out=out))
        res.append({"op": "nt", "m": m, "n": n, "k": k, "bn": bn, "ms": ms, "tflops": 2 * m * n * k / ms / 1e9,
                    "torch_ms": ms_t, "torch_tflops": 2 * m * n * k / ms_t / 1e9})
        print(res[-1], flush=True)
    # epilogue variants at the fc1 shape
    m, n, k = 28416, 4096, 1024
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    for name, kw in [("plain", {}), ("bias", dict(bias=bias)), ("bias+gelu", dict(bias=bias, act=1)),
                     ("bias+gelu+preact", dict(bias=bias, act=1, preact=pre)), ("dgelu", dict(dgelu_u=pre)),
                     ("residual", dict(residual=pre))]:
        ms = bench(lambda: gemm.gemm_nt(a, w, out=out, **kw))
        res.append({"op": "nt-epi " + name, "m": m, "n": n, "k": k, "ms": ms, "tflops": 2 * m * n * k / ms / 1e9})
        print(res[-1], flush=True)
    del a, w, out, pre
    for (r, m, n, bn) in [(54528, 1024, 1024, 256), (54528, 4096, 1024, 256), (54528, 1024, 4096, 256)]:
        a = torch.randn(r, m, device="cuda").bfloat16()
        b = torch.randn(r, n, device="cuda").bfloat16()
        out = torch.zeros(m, n, device="cuda")
        ms = bench(lambda: gemm.gemm_tn(a, b, out, block_n=bn))
        res.append({"op": "tn", "r": r, "m": m, "n": n, "bn": bn, "ms": ms, "tflops": 2 * m * n * r / ms / 1e9})
        print(res[-1], flush=True)
    # grouped conv (positional encoder shape): 24 clones x 2000 x 1024, 16 groups, 19 taps
    x = torch.randn(24, 2000, 1024, device="cuda").bfloat16()
    w = (torch.randn(1024, 19 * 64, device="cuda") * 0.05).bfloat16()
    out = torch.empty_like(x)
    ms = bench(lambda: gemm.conv_nt(x, w, taps=19, pad=9, groups=16, out=out))
    fl = 2 * 24 * 2000 * 1024 * 64 * 19
    res.append({"op": "conv_nt", "ms": ms, "tflops": fl / ms / 1e9})
    print(res[-1], flush=True)
    for mode in (0,):
        try:
            o2 = gemm.conv_slab(x, w, taps=19, pad=9, groups=16)
            err = ((o2.float() - out.float()).norm() / out.float().norm()).item()
            ms = bench(lambda: gemm.conv_slab(x, w, taps=19, pad=9, groups=16, out=o2))
            res.append({"op": "conv_slab", "ms": ms, "tflops": fl / ms / 1e9, "rel_vs_conv_nt": err})
            print(res[-1], flush=True)
        except Exception as ex:
            print("conv_slab failed", mode, ex, flush=True)
    dw = torch.zeros(16 * 19 * 64, 64, device="cuda")
    ms = bench(lambda: gemm.conv_wgrad_tn(x, x, dw, taps=19, pad=9, groups=16))
    res.append({"op": "conv_wgrad_tn", "ms": ms, "tflops": fl / ms / 1e9})
    print(res[-1], flush=True)
    ms = bench(lambda: gemm.conv_slab_wgrad(x, x, dw, taps=19, pad=9, groups=16))
    res.append({"op": "conv_slab_wgrad", "ms": ms, "tflops": fl / ms / 1e9})
    print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
