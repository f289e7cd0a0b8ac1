"""GPU parity of the tcgen05 GEMM against fp32 torch math on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _randn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize(
    "m,n,k,bn",
    [(128, 128, 64, 128), (256, 256, 512, 128), (1000, 384, 264, 128), (777, 1024, 1024, 256), (4096, 64, 128, 64),
     (300, 200, 72, 256), (2048, 3072, 1024, 256)],
)
def test_gemm_nt_plain(m, n, k, bn):
    from animal2vec_b200 import gemm

    a, w = _randn(m, k, seed=1), _randn(n, k, seed=2)
    ref = a.float() @ w.float().t()
    out = gemm.gemm_nt(a, w, out_dtype=torch.float32, block_n=bn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
    out16 = gemm.gemm_nt(a, w, block_n=bn)
    assert _rel(out16, ref) < 4e-3


def test_gemm_nt_epilogues():
    from animal2vec_b200 import gemm

    m, n, k = 515, 320, 256
    a, w = _randn(m, k, seed=3), _randn(n, k, scale=0.1, seed=4)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda")
    u_ref = a.float() @ w.float().t() * 0.5 + bias
    pre = torch.empty(m, n, device="cuda")
    out = gemm.gemm_nt(a, w, out_dtype=torch.float32, bias=bias, act=1, preact=pre, residual=res, alpha=0.5)
    assert _rel(pre, u_ref) < 1e-5
    assert _rel(out, F.gelu(u_ref) + res) < 1e-5
    # dgelu epilogue: out = (a w^T) * gelu'(u)
    u = torch.randn(m, n, device="cuda")
    out2 = gemm.gemm_nt(a, w, out_dtype=torch.float32, dgelu_u=u)
    uu = u.clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    assert _rel(out2, (a.float() @ w.float().t()) * uu.grad) < 1e-5
    # the compile-time bf16 epilogues (8-warp GELU variants with the staged bias vector; bias-only; residual)
    for (mm, nn, kk) in ((515, 320, 256), (1000, 4096, 1024)):
        a2, w2 = _randn(mm, kk, seed=5), _randn(nn, kk, scale=0.05, seed=6)
        b2 = torch.randn(nn, device="cuda")
        r2 = torch.randn(mm, nn, device="cuda").bfloat16()
        u2 = a2.float() @ w2.float().t() + b2
        pre16 = torch.empty(mm, nn, device="cuda", dtype=torch.bfloat16)
        o1 = gemm.gemm_nt(a2, w2, out_dtype=torch.bfloat16, bias=b2, act=1, preact=pre16)
        assert _rel(pre16, u2) < 4e-3 and _rel(o1, F.gelu(u2)) < 5e-3
        o2 = gemm.gemm_nt(a2, w2, out_dtype=torch.bfloat16, bias=b2, act=1)
        assert torch.equal(o1, o2)
        assert (o1.float() - F.gelu(u2)).abs().max().item() < 2e-3 + 8e-3 * F.gelu(u2).abs().max().item()
        assert _rel(gemm.gemm_nt(a2, w2, out_dtype=torch.bfloat16, bias=b2), u2) < 4e-3
        assert _rel(gemm.gemm_nt(a2, w2, out_dtype=torch.bfloat16, residual=r2), u2 - b2 + r2.float()) < 4e-3
    # accumulate
    acc = torch.ones(m, n, device="cuda")
    gemm.gemm_nt(a, w, out=acc, accumulate=True)
    assert _rel(acc, 1 + a.float() @ w.float().t()) < 1e-5


@pytest.mark.parametrize("bsz,t,cg,ng,groups,taps", [(2, 300, 64, 64, 4, 19), (3, 130, 64, 64, 2, 7), (1, 257, 128, 192, 1, 3),
                                                    (2, 2000, 64, 64, 16, 19)])
def test_conv_nt(bsz, t, cg, ng, groups, taps):
    from animal2vec_b200 import gemm

    pad = taps // 2
    x = _randn(bsz, t, groups * cg, seed=5)
    wt = _randn(groups * ng, cg, taps, scale=0.05, seed=6)  # torch layout (out, in/groups, k)
    bias = torch.randn(groups * ng, device="cuda")
    ref = F.conv1d(x.float().transpose(1, 2), wt.float(), bias, padding=pad, groups=groups).transpose(1, 2)
    w = wt.permute(0, 2, 1).reshape(groups * ng, taps * cg).contiguous()
    out = gemm.conv_nt(x, w, taps=taps, pad=pad, groups=groups, out_dtype=torch.float32, bias=bias)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


@pytest.mark.parametrize("r,m,n,bn,ks", [(64, 128, 128, 128, 1), (1000, 256, 320, 128, None), (5000, 1024, 512, 256, None),
                                         (333, 104, 64, 64, 3), (4096, 128, 4096, 256, 1)])
def test_gemm_tn(r, m, n, bn, ks):
    from animal2vec_b200 import gemm

    a, b = _randn(r, m, seed=7), _randn(r, n, seed=8)
    ref = a.float().t() @ b.float()
    out = torch.zeros(m, n, device="cuda")
    gemm.gemm_tn(a, b, out, block_n=bn, k_splits=ks)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
    gemm.gemm_tn(a, b, out, block_n=bn, k_splits=ks)  # accumulates
    assert _rel(out, 2 * ref) < 1e-5


@pytest.mark.parametrize("bsz,t,cg,ng,groups,taps", [(2, 300, 64, 64, 4, 19), (3, 130, 64, 64, 2, 7), (2, 257, 128, 128, 1, 3),
                                                    (2, 200, 512, 512, 1, 3), (1, 333, 128, 320, 1, 3)])
def test_conv_wgrad_tn(bsz, t, cg, ng, groups, taps):
    from animal2vec_b200 import gemm

    pad = taps // 2
    x = _randn(bsz, t, groups * cg, seed=9)
    dy = _randn(bsz, t, groups * ng, seed=10)
    xt = x.float().transpose(1, 2).requires_grad_(False)
    wt = torch.zeros(groups * ng, cg, taps, device="cuda", requires_grad=True)
    y = F.conv1d(xt, wt, None, padding=pad, groups=groups)
    y.backward(dy.float().transpose(1, 2))
    # transposed layout: out[(g*taps + j)*cg + c, n] = dW[g*ng + n, c, j]
    ref = wt.grad.view(groups, ng, cg, taps).permute(0, 3, 2, 1).reshape(groups * taps * cg, ng)
    out = torch.zeros(groups * taps * cg, ng, device="cuda")
    gemm.conv_wgrad_tn(dy, x, out, taps=taps, pad=pad, groups=groups)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


@pytest.mark.parametrize("bsz,t,ng,groups,taps,pad", [(2, 300, 64, 4, 19, 9), (3, 130, 64, 2, 7, 3), (1, 2000, 64, 16, 19, 9),
                                                       (2, 257, 48, 3, 7, 3), (2, 512, 64, 1, 3, 1), (1, 700, 64, 2, 19, 9)])
def test_conv_slab(bsz, t, ng, groups, taps, pad):
    """Slab implicit GEMM (rows loaded once, tap = shifted shared-memory descriptor) against torch conv1d."""
    from animal2vec_b200 import gemm

    cg = 64
    x = _randn(bsz, t, groups * cg, seed=11)
    wt = _randn(groups * ng, cg, taps, scale=0.05, seed=12)
    bias = torch.randn(groups * ng, device="cuda")
    ref = F.conv1d(x.float().transpose(1, 2), wt.float(), bias, padding=pad, groups=groups).transpose(1, 2)
    w = wt.permute(0, 2, 1).reshape(groups * ng, taps * cg).contiguous()
    assert gemm.conv_slab_ok(x, w, taps, groups)
    out = gemm.conv_slab(x, w, taps=taps, pad=pad, groups=groups, out_dtype=torch.float32, bias=bias)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
    out16 = gemm.conv_slab(x, w, taps=taps, pad=pad, groups=groups, bias=bias)
    assert _rel(out16, ref) < 5e-3


@pytest.mark.parametrize("real", [16, 48])
def test_conv_slab_skips_the_zero_padded_channels_of_a_group(real):
    """Group-padded layout (the decoder: 48 real channels in 64-wide groups, zero weights over the padding): with
    ``x_real_cols`` the kernel skips the K steps over the padding -- same result whatever the padding columns hold."""
    from animal2vec_b200 import gemm

    bsz, t, groups, taps, pad, cg = 2, 300, 4, 7, 3, 64
    x = _randn(bsz, t, groups * cg, seed=21)
    wt = _randn(groups * 64, cg, taps, scale=0.05, seed=22)
    wt.view(groups, 64, cg, taps)[:, :, real:] = 0
    ref = F.conv1d(x.float().transpose(1, 2), wt.float(), None, padding=pad, groups=groups).transpose(1, 2)
    w = wt.permute(0, 2, 1).reshape(groups * 64, taps * cg).contiguous()
    full = gemm.conv_slab(x, w, taps=taps, pad=pad, groups=groups, out_dtype=torch.float32)
    xg = x.clone()
    xg.view(bsz, t, groups, cg)[..., real:] = 1.0e4  # finite garbage in the padding columns
    skip = gemm.conv_slab(xg, w, taps=taps, pad=pad, groups=groups, out_dtype=torch.float32, x_real_cols=real)
    assert _rel(full, ref) < 1e-5 and _rel(skip, ref) < 1e-5, (_rel(full, ref), _rel(skip, ref))


@pytest.mark.parametrize("bsz,t,ng,groups,taps,pad", [(2, 300, 64, 4, 19, 9), (3, 130, 64, 2, 7, 3), (2, 2000, 64, 16, 19, 9),
                                                       (2, 257, 48, 3, 7, 3), (5, 64, 64, 1, 3, 1), (1, 700, 64, 2, 25, 12)])
def test_conv_slab_wgrad(bsz, t, ng, groups, taps, pad):
    from animal2vec_b200 import gemm

    cg = 64
    x = _randn(bsz, t, groups * cg, seed=13)
    dy = _randn(bsz, t, groups * ng, seed=14)
    wt = torch.zeros(groups * ng, cg, taps, device="cuda", requires_grad=True)
    y = F.conv1d(x.float().transpose(1, 2), wt, None, padding=pad, groups=groups)
    y.backward(dy.float().transpose(1, 2))
    ref = wt.grad.view(groups, ng, cg, taps).permute(0, 3, 2, 1).reshape(groups * taps * cg, ng)
    out = torch.zeros(groups * taps * cg, ng, device="cuda")
    gemm.conv_slab_wgrad(dy, x, out, taps=taps, pad=pad, groups=groups)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
    gemm.conv_slab_wgrad(dy, x, out, taps=taps, pad=pad, groups=groups)  # accumulates
    assert _rel(out, 2 * ref) < 1e-5


@pytest.mark.parametrize("m,n,k", [(600, 256, 128), (1500, 512, 320), (4100, 4096, 1024)])
def test_cta_pair_gemm_fused_gelu_backward_and_column_sums(m, n, k):
    """Epilogue of the fc2 data-gradient product: result * GELU'(u) with the column sums of the stored result (= the fc1
    bias gradient) accumulated on top of what the buffer holds; ragged M (rows beyond M must not reach the sums)."""
    from animal2vec_b200 import gemm

    a, w = _randn(m, k, seed=31), _randn(n, k, scale=0.05, seed=32)
    u = _randn(m, n, seed=33)
    uf = u.float().requires_grad_(True)
    F.gelu(uf, approximate="tanh").sum().backward()
    ref = (a.float() @ w.float().t()) * uf.grad
    acc = torch.full((n,), 0.5, device="cuda")
    assert gemm.pair_shape_ok(m, n, k)
    out = gemm.gemm_nt(a, w, dgelu_u=u, colsum=acc)
    assert _rel(out, ref) < 5e-3, _rel(out, ref)
    assert _rel(acc, 0.5 + ref.sum(0)) < 2e-3, _rel(acc, 0.5 + ref.sum(0))  # fp32 sums of the unrounded products
    plain = gemm.gemm_nt(a, w, dgelu_u=u)
    assert torch.equal(plain, out)


@pytest.mark.parametrize("m,n,k", [(513, 256, 64), (1500, 512, 320), (4000, 1024, 1024)])
def test_cta_pair_gemm_shapes_and_epilogues(m, n, k):
    """Shapes the CTA-pair (cta_group::2) kernels take (N a multiple of 256, K of 64): ragged M (the second CTA of the last
    pair partly or wholly out of range), every compile-time epilogue, and the TN weight-gradient product."""
    from animal2vec_b200 import gemm

    a, w = _randn(m, k, seed=11), _randn(n, k, scale=0.05, seed=12)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda").bfloat16()
    u = a.float() @ w.float().t()
    pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    assert _rel(gemm.gemm_nt(a, w), u) < 4e-3
    assert _rel(gemm.gemm_nt(a, w, bias=bias, alpha=0.5), 0.5 * u + bias) < 4e-3
    assert _rel(gemm.gemm_nt(a, w, bias=bias, act=1), F.gelu(u + bias)) < 5e-3
    assert _rel(gemm.gemm_nt(a, w, bias=bias, act=1, preact=pre), F.gelu(u + bias)) < 5e-3 and _rel(pre, u + bias) < 4e-3
    assert _rel(gemm.gemm_nt(a, w, residual=res), u + res.float()) < 4e-3
    # TN: out (n_a, n_b) += a^T b over the m rows
    b2 = _randn(m, 256, seed=13)
    a2 = _randn(m, n, seed=14)
    out = torch.ones(n, 256, device="cuda")
    gemm.gemm_tn(a2, b2, out)
    assert _rel(out, 1 + a2.float().t() @ b2.float()) < 1e-5


@pytest.mark.parametrize("bsz,t,c,c_real,n,k,s,pad", [(2, 4000, 128, 127, 512, 10, 5, 3), (3, 1000, 512, 512, 512, 3, 2, 1),
                                                        (2, 640, 64, 64, 64, 3, 2, 1), (1, 80000, 128, 127, 512, 10, 5, 3)])
def test_strided_conv_without_im2col(bsz, t, c, c_real, n, k, s, pad):
    """Strided Conv1d forward, data gradient and weight gradient through the (B, T/s, s*C) view addressing of the tap-loop
    GEMM (no im2col / col2im buffers) against F.conv1d autograd; c_real < c exercises the zero-padded input channel of
    the sinc layer (127 of 128)."""
    from animal2vec_b200 import gemm, ops
    from animal2vec_b200 import params as P

    w = _randn(n, c_real, k, scale=0.05, seed=1).float()  # fp32 master holding bf16-representable values
    x = _randn(bsz, t, c, seed=2)
    x[..., c_real:] = 0
    pk_f = P.pack_conv_fwd("w", 1, n, c_real, k, cgp=c)
    pk_d = P.pack_conv_dgrad_strided("w", n, c_real, k, s, pad, c)
    wf = P.materialize(pk_f, w, False)
    wd = P.materialize(pk_d, w, False)
    xr = x.float().clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = torch.nn.functional.conv1d(xr[..., :c_real].transpose(1, 2), wr, stride=s, padding=pad).transpose(1, 2)
    assert ref.shape[1] == t // s
    y = gemm.strided_conv_nt(x, wf, taps=k, stride=s, pad=pad)
    assert _rel(y, ref) < 6e-3, _rel(y, ref)
    dy = _randn(bsz, t // s, n, seed=3)
    ref.backward(dy.float())
    dx = gemm.strided_conv_dgrad(dy, wd, c=c, taps_per_block=-(-k // s), stride=s, pad=pad)
    assert _rel(dx[..., :c_real], xr.grad[..., :c_real]) < 6e-3, _rel(dx[..., :c_real], xr.grad[..., :c_real])
    assert float(dx[..., c_real:].abs().sum()) == 0.0
    gw = torch.zeros(pk_f.gt_shape, device="cuda")
    gemm.strided_conv_wgrad_tn(dy, x, gw, taps=k, stride=s, pad=pad)
    got = torch.zeros_like(w)
    P.unpack_grad(pk_f, gw, got, transposed=True)
    assert _rel(got, wr.grad) < 6e-3, _rel(got, wr.grad)


@pytest.mark.parametrize("rows,t,tk,groups,taps", [(6, 300, 23, 2, 19), (24, 2000, 148, 16, 19), (3, 130, 130, 4, 7)])
def test_gathered_conv_on_kept_rows(rows, t, tk, groups, taps):
    """The last positional-conv layer evaluated only at the kept positions: neighbourhood row map + row gather +
    tap-blocked grouped GEMM (forward) and its weight gradient, against F.conv1d over the full sequence read at the kept
    rows / autograd with a gradient that is zero elsewhere."""
    from animal2vec_b200 import gemm, ops
    from animal2vec_b200 import params as P

    c = groups * 64
    pad = taps // 2
    x = _randn(rows, t, c, seed=1)
    w = _randn(c, 64, taps, scale=0.05, seed=2).float()
    bias = _randn(c, seed=3).float()
    g = torch.Generator().manual_seed(4)
    ids = torch.stack([torch.randperm(t, generator=g)[:tk].sort().values for _ in range(rows)]).to(torch.int32).cuda()
    pk = P.pack_conv_fwd("w", groups, 64, 64, taps)
    wf = P.materialize(pk, w, False)
    nidx = ops.neigh_index(ids, t, taps, pad)
    ref_idx = (torch.arange(rows, device="cuda")[:, None, None] * t + ids.long()[:, :, None]
               + torch.arange(taps, device="cuda")[None, None, :] - pad)
    inside = (ids.long()[:, :, None] + torch.arange(taps, device="cuda") - pad >= 0) & \
             (ids.long()[:, :, None] + torch.arange(taps, device="cuda") - pad < t)
    assert torch.equal(nidx.view(rows, tk, taps).long(), torch.where(inside, ref_idx, torch.full_like(ref_idx, -1)))
    xg = ops.row_gather(x.view(rows * t, c), nidx, nidx.numel()).view(rows * tk, taps, c)
    y = gemm.gathered_conv_nt(xg, wf, taps=taps, groups=groups, bias=bias)
    xr = x.float()
    wr = w.clone().requires_grad_(True)
    full = F.conv1d(xr.transpose(1, 2), wr, bias, padding=pad, groups=groups).transpose(1, 2)  # (rows, t, c)
    ref = torch.gather(full, 1, ids.long().unsqueeze(-1).expand(-1, -1, c)).reshape(rows * tk, c)
    assert _rel(y, ref) < 6e-3, _rel(y, ref)
    dy = _randn(rows * tk, c, seed=5)
    ref.backward(dy.float())
    gw = torch.zeros(pk.gt_shape, device="cuda")
    gemm.gathered_conv_wgrad_tn(dy, xg, gw, taps=taps, groups=groups)
    got = torch.zeros_like(w)
    P.unpack_grad(pk, gw, got, transposed=True)
    assert _rel(got, wr.grad) < 6e-3, _rel(got, wr.grad)
