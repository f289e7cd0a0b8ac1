"""GPU parity of every non-GEMM kernel against plain fp32 torch math of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


def _pswish(n, al, be):
    return n * al * torch.sigmoid(be * n)


# --------------------------------------------------------------------------------------- row LN
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize(
    "rows,c,gw,gr,act,affine,two,post",
    [(37, 128, 128, 127, 2, True, False, False), (1000, 512, 512, 512, 1, True, False, False),
     (513, 1024, 1024, 1024, 0, True, True, False), (200, 1024, 64, 48, 1, False, False, True),
     (64, 768, 768, 768, 1, False, False, False), (33, 64, 64, 64, 0, True, True, False),
     (2500, 768, 768, 768, 0, True, True, False), (1301, 512, 512, 512, 0, True, False, False),
     (3000, 1024, 1024, 1024, 0, True, False, False)],
)
def test_rowln(dtype, rows, c, gw, gr, act, affine, two, post):
    from animal2vec_b200 import ops

    creal = c // gw * gr
    real = (torch.arange(c, device="cuda") % gw) < gr
    a = torch.randn(rows, c, device="cuda", generator=_g(1)) * real
    b = torch.randn(rows, c, device="cuda", generator=_g(2)) * real if two else None
    po = torch.randn(rows, c, device="cuda", generator=_g(3)) * real if post else None
    gamma = torch.randn(creal, device="cuda", generator=_g(4)) if affine else None
    beta = torch.randn(creal, device="cuda", generator=_g(5)) if affine else None
    al = (torch.randn(creal, device="cuda", generator=_g(6)) + 2) if act == 2 else None
    be = torch.randn(creal, device="cuda", generator=_g(7)) if act == 2 else None
    dy = torch.randn(rows, c, device="cuda", generator=_g(8)) * real
    a, dy = a.to(dtype), dy.to(dtype)
    b = b.to(dtype) if two else None
    po = po.to(dtype) if post else None

    # torch reference on the real channels
    ar = a.float()[:, real].clone().requires_grad_(True)
    br = b.float()[:, real].clone().requires_grad_(True) if two else None
    params = [t.clone().requires_grad_(True) if t is not None else None for t in (gamma, beta, al, be)]
    z = ar + (br if two else 0)
    n = F.layer_norm(z, (creal,), params[0], params[1], 1e-5)
    y = F.gelu(n) if act == 1 else (_pswish(n, params[2], params[3]) if act == 2 else n)
    if post:
        y = y + po.float()[:, real]
    y.backward(dy.float()[:, real])

    cfg = ops.RowLnCfg(c, 1e-5, act=act, group_width=gw, group_real=gr)
    out, mean, rstd = ops.rowln_fwd(cfg, a, b, gamma, beta, al, be, po)
    tol = 1e-5 if dtype == torch.float32 else 8e-3
    assert _rel(out.float()[:, real], y) < tol
    assert out.float()[:, ~real].abs().sum().item() == 0
    dg = torch.zeros(creal, device="cuda") if affine else None
    dbt = torch.zeros(creal, device="cuda") if affine else None
    dal = torch.zeros(creal, device="cuda") if act == 2 else None
    dbe = torch.zeros(creal, device="cuda") if act == 2 else None
    da, db = ops.rowln_bwd(cfg, dy, a, b, gamma, beta, al, be, mean, rstd, dgamma=dg, dbeta=dbt, dact_alpha=dal,
                           dact_beta=dbe)
    assert _rel(da.float()[:, real], ar.grad) < tol
    if two:
        assert _rel(db.float()[:, real], br.grad) < tol
    if affine:
        assert _rel(dg, params[0].grad) < max(tol, 2e-5) and _rel(dbt, params[1].grad) < max(tol, 2e-5)
    if act == 2:
        assert _rel(dal, params[2].grad) < max(tol, 2e-5) and _rel(dbe, params[3].grad) < max(tol, 2e-5)


def test_rowln_dropout_consistency():
    """fwd/bwd regenerate the same dropout masks; keep rate is ~1-p; scaling is 1/(1-p)."""
    from animal2vec_b200 import ops

    rows, c = 512, 1024
    a = torch.zeros(rows, c, device="cuda")
    b = torch.ones(rows, c, device="cuda")
    cfg = ops.RowLnCfg(c, 1e-5, drop_b=0.1)
    # LN of a constant-with-holes row: use the backward's db to read the mask
    y, mean, rstd = ops.rowln_fwd(cfg, a, b, seed_b=1234)
    dy = torch.randn(rows, c, device="cuda", generator=_g(1))
    da, db = ops.rowln_bwd(cfg, dy, a, b, None, None, None, None, mean, rstd, seed_b=1234)
    keep = db != 0
    frac = keep.float().mean().item()
    assert abs(frac - 0.9) < 0.01, frac
    assert torch.allclose(db[keep], da[keep] / 0.9, rtol=1e-5, atol=1e-7)
    # forward used the same mask: z = keep/0.9 -> y = LN(z); check against torch
    z = keep.float() / 0.9
    assert _rel(y, F.layer_norm(z, (c,), None, None, 1e-5)) < 1e-4
    # output dropout
    cfg2 = ops.RowLnCfg(c, 1e-5, drop_out=0.25)
    x = torch.randn(rows, c, device="cuda", generator=_g(2))
    y2, m2, r2 = ops.rowln_fwd(cfg2, x, seed_out=99)
    ref = F.layer_norm(x, (c,), None, None, 1e-5)
    k2 = y2 != 0
    assert abs(k2.float().mean().item() - 0.75) < 0.01
    assert torch.allclose(y2[k2], ref[k2] / 0.75, rtol=1e-4, atol=1e-5)
    y3, _, _ = ops.rowln_fwd(cfg2, x, seed_out=99, training=False)
    assert _rel(y3, ref) < 1e-5


@pytest.mark.parametrize("c", [512, 768, 1024])
def test_rowln_residual_bf16_dropout_matches_generic_path(c):
    """The specialised bf16 residual-norm kernels (LN(a + drop(b)) * gamma + beta, fused bias-gradient column
    sums) against the generic fp32 kernels with the SAME dropout seed (identical keep masks by construction)."""
    from animal2vec_b200 import ops

    rows = 2999
    a = torch.randn(rows, c, device="cuda", generator=_g(1)).bfloat16()
    b = torch.randn(rows, c, device="cuda", generator=_g(2)).bfloat16()
    dy = torch.randn(rows, c, device="cuda", generator=_g(3)).bfloat16()
    gamma = torch.randn(c, device="cuda", generator=_g(4))
    beta = torch.randn(c, device="cuda", generator=_g(5))
    cfg = ops.RowLnCfg(c, 1e-5, drop_b=0.1)
    res = {}
    for name, cast in (("bf16", lambda t: t), ("fp32", lambda t: t.float())):
        y, m, r = ops.rowln_fwd(cfg, cast(a), cast(b), gamma, beta, seed_b=77)
        dg, dbt, dbias = (torch.zeros(c, device="cuda") for _ in range(3))
        da, db = ops.rowln_bwd(cfg, cast(dy), cast(a), cast(b), gamma, beta, None, None, m, r, seed_b=77, dgamma=dg,
                               dbeta=dbt, dbias_b=dbias)
        res[name] = (y.float(), da.float(), db.float(), dg, dbt, dbias, m, r)
    for i, tol in enumerate((8e-3, 8e-3, 8e-3, 8e-3, 8e-3, 2e-2, 1e-4, 1e-4)):
        assert _rel(res["bf16"][i], res["fp32"][i]) < tol, i
    assert torch.equal(res["bf16"][2] != 0, res["fp32"][2] != 0)  # identical dropout masks
    assert _rel(res["bf16"][5], res["bf16"][2].sum(0)) < 1e-2      # fused column sums = sums of the stored db
    assert abs((res["bf16"][2] != 0).float().mean().item() - 0.9) < 0.01


# --------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, batch, seq, heads, pos, slopes, scale):
    d = heads * 64
    q, k, v = qkv.float().view(batch, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q * 64 ** -0.5) @ k.transpose(-1, -2)
    p = pos.float() if pos is not None else torch.arange(seq, device=qkv.device).float().expand(batch, seq)
    dist = (p[:, :, None] - p[:, None, :]).abs()
    coef = slopes * scale.clamp_min(0).view(-1)
    s = s - coef.view(1, heads, 1, 1) * dist[:, None]
    a = s.softmax(-1)
    return (a @ v).transpose(1, 2).reshape(batch, seq, d)


@pytest.mark.parametrize("dtype,batch,seq,heads,with_pos",
                         [(torch.bfloat16, 2, 300, 2, False), (torch.bfloat16, 3, 142, 4, True),
                          (torch.bfloat16, 1, 2000, 2, False), (torch.float32, 2, 77, 2, True),
                          (torch.bfloat16, 2, 128, 1, True), (torch.bfloat16, 2, 129, 1, False)])
def test_attention_fwd_bwd(dtype, batch, seq, heads, with_pos):
    from animal2vec_b200 import ops

    d = heads * 64
    qkv = torch.randn(batch, seq, 3 * d, device="cuda", generator=_g(1)).to(dtype)
    pos = None
    if with_pos:
        pos = torch.stack([torch.randperm(2000, device="cuda", generator=_g(10 + i))[:seq].sort().values
                           for i in range(batch)]).to(torch.int32).contiguous()
    slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(heads)], device="cuda")
    scale = torch.rand(heads, device="cuda", generator=_g(2)) + 0.5
    qr = qkv.float().clone().requires_grad_(True)
    sr = scale.clone().requires_grad_(True)
    ref = _attn_ref(qr, batch, seq, heads, pos, slopes, sr)
    dout = torch.randn(batch, seq, d, device="cuda", generator=_g(3)).to(dtype)
    ref.backward(dout.float())

    out, lse = ops.attn_fwd(qkv, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert _rel(out, ref) < tol, _rel(out, ref)
    if True:  # every length: resident kernel up to 160 tokens, tiled kernel beyond (bf16); fp32 validation kernels
        dsc = torch.zeros(heads, device="cuda")
        dqkv = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale,
                            dalibi_scale=dsc)
        gt = 5e-5 if dtype == torch.float32 else 2e-2
        assert _rel(dqkv, qr.grad) < gt, _rel(dqkv, qr.grad)
        assert _rel(dsc, sr.grad) < max(gt, 1e-3), (dsc, sr.grad)


@pytest.mark.parametrize("batch,seq,heads", [(5, 148, 16), (300, 142, 4), (2, 97, 2), (1, 129, 3)])
def test_attention_backward_fused_qkv_bias_gradient(batch, seq, heads):
    """The resident backward accumulates the column sums of dqkv (= the qkv bias gradient) per CTA in shared memory;
    must equal the separate column-sum pass over the stored dqkv, for grids smaller and larger than the SM count."""
    from animal2vec_b200 import ops

    d = heads * 64
    qkv = torch.randn(batch, seq, 3 * d, device="cuda", generator=_g(1)).bfloat16()
    pos = torch.stack([torch.randperm(2000, device="cuda", generator=_g(10 + i % 7))[:seq].sort().values
                       for i in range(batch)]).to(torch.int32).contiguous()
    slopes = torch.tensor([2.0 ** (-8.0 * (h + 1) / heads) for h in range(heads)], device="cuda")
    scale = torch.ones(heads, device="cuda")
    out, lse = ops.attn_fwd(qkv, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=5)
    dout = torch.randn(batch, seq, d, device="cuda", generator=_g(3)).bfloat16()
    acc = torch.full((3 * d,), 0.25, device="cuda")  # accumulates on top of what is there
    dqkv = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1,
                        seed=5, dqkv_colsum=acc)
    plain = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=0.1, seed=5)
    assert torch.equal(dqkv, plain)  # the work distribution (head-major vs interleaved) does not change the result
    want = dqkv.float().view(-1, 3 * d).sum(0) + 0.25
    assert _rel(acc, want) < 3e-3, _rel(acc, want)  # fp32 pre-rounding sums vs sums of the bf16-rounded rows


@pytest.mark.parametrize("batch,seq,heads,with_pos,drop_p",
                         [(3, 142, 4, True, 0.0), (2, 128, 2, False, 0.0), (2, 129, 2, True, 0.1), (2, 300, 2, False, 0.0),
                          (2, 852, 4, True, 0.0), (1, 1000, 2, True, 0.2), (2, 2000, 16, False, 0.0), (1, 2100, 3, False, 0.1)])
def test_attention_backward_tiled_any_length(batch, seq, heads, with_pos, drop_p):
    """The tiled backward (key tiles resident, query tiles streamed, dQ through fp32 atomics; finetune path and 48 kHz
    student) against fp32 torch autograd: ragged lengths, token positions, the ALiBi window on contiguous sequences,
    and attention dropout with the kernel's own keep mask (forward and backward regenerate the same hashes). Forced
    (algo=2) also on the short lengths the resident kernel normally takes, where both kernels must agree."""
    from animal2vec_b200 import ops

    d = heads * 64
    seed = 0x1234ABCD99
    qkv = torch.randn(batch, seq, 3 * d, device="cuda", generator=_g(1)).bfloat16()
    pos = None
    if with_pos:
        pos = torch.stack([torch.randperm(12000, device="cuda", generator=_g(10 + i))[:seq].sort().values
                           for i in range(batch)]).to(torch.int32).contiguous()
    slopes = torch.tensor([2.0 ** (-8.0 * (h + 1) / heads) for h in range(heads)], device="cuda")
    scale = torch.rand(heads, device="cuda", generator=_g(2)) + 0.5
    if heads > 2:
        scale[1] = 0.0  # a head without ALiBi: no window, clamp gate on its scale gradient
    dout = torch.randn(batch, seq, d, device="cuda", generator=_g(3)).bfloat16()
    out, lse = ops.attn_fwd(qkv, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale, drop_p=drop_p, seed=seed,
                            skip_far_keys=True)

    qr = qkv.float().clone().requires_grad_(True)
    sr = scale.clone().requires_grad_(True)
    q, k, v = qr.view(batch, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q * 64 ** -0.5) @ k.transpose(-1, -2)
    p_ = pos.float() if pos is not None else torch.arange(seq, device="cuda").float().expand(batch, seq)
    dist = (p_[:, :, None] - p_[:, None, :]).abs()
    s = s - (slopes * sr.clamp_min(0)).view(1, heads, 1, 1) * dist[:, None]
    a = s.softmax(-1)
    if drop_p > 0:
        keep = torch.zeros(batch, heads, seq, seq, device="cuda", dtype=torch.bool)
        for c0 in range(0, seq, 64):  # read the mask out through the forward's linearity in V (uniform attention)
            z = torch.zeros(batch, seq, 3 * d, device="cuda", dtype=torch.bfloat16)
            n = min(64, seq - c0)
            z[:, c0:c0 + n, 2 * d:] = torch.eye(64, device="cuda", dtype=torch.bfloat16)[:n].repeat(1, heads)
            o, _ = ops.attn_fwd(z, batch, seq, heads, pos=pos, drop_p=drop_p, seed=seed)
            keep[:, :, :, c0:c0 + n] = o.float().view(batch, seq, heads, 64).permute(0, 2, 1, 3)[..., :n] > 0.5 / seq
        a = a * keep.float() / (1 - drop_p)
    ref = (a @ v).transpose(1, 2).reshape(batch, seq, d)
    ref.backward(dout.float())
    assert _rel(out, ref) < 1e-2

    dsc = torch.zeros(heads, device="cuda")
    dqkv = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale,
                        dalibi_scale=dsc, drop_p=drop_p, seed=seed, algo=2)
    for part, name in ((slice(0, d), "dq"), (slice(d, 2 * d), "dk"), (slice(2 * d, 3 * d), "dv")):
        e = _rel(dqkv[..., part], qr.grad[..., part])
        assert e < 2e-2, (name, e)
    assert _rel(dsc, sr.grad) < 2e-2, (dsc, sr.grad)
    if seq <= 160:
        dsc1 = torch.zeros(heads, device="cuda")
        dq1 = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, pos=pos, slopes=slopes, alibi_scale=scale,
                           dalibi_scale=dsc1, drop_p=drop_p, seed=seed, algo=1)
        assert _rel(dqkv, dq1) < 5e-3 and _rel(dsc, dsc1) < 5e-3
    if pos is None and seq > 256:  # the window changes nothing measurable
        dq_full = ops.attn_bwd(dout, qkv, out, lse, batch, seq, heads, slopes=slopes, alibi_scale=scale, drop_p=drop_p,
                               seed=seed, algo=2, skip_far_keys=False)
        assert _rel(dqkv, dq_full) < 1e-3, _rel(dqkv, dq_full)


@pytest.mark.parametrize("qk_scale,seq", [(1.0, 2000), (0.3, 2000), (6.0, 2000), (1.0, 1111), (1.0, 2500)])
def test_attention_alibi_locality_skip_is_exact(qk_scale, seq):
    """Full-length (teacher) attention with the ALiBi key-tile window: identical to the full sweep and to fp32 torch
    math for small, unit and large q/k norms (the window widens with max|q| max|k|; with large norms nothing is
    skipped), for ragged lengths, with the model's 16 slopes and a zero-scale head (no ALiBi -> no skipping)."""
    from animal2vec_b200 import ops

    batch, heads = 2, 16
    d = heads * 64
    qkv = (torch.randn(batch, seq, 3 * d, device="cuda", generator=_g(1)) * qk_scale).bfloat16()
    qkv[..., 2 * d:] = torch.randn(batch, seq, d, device="cuda", generator=_g(2)).bfloat16()
    slopes = torch.tensor([2.0 ** (-0.5 * (h + 1)) for h in range(heads)], device="cuda")
    scale = torch.rand(heads, device="cuda", generator=_g(3)) + 0.5
    scale[5] = 0.0
    full, lse_full = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale)
    fast, lse_fast = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale, skip_far_keys=True)
    # P is rounded to bf16 relative to a different reference exponent in the two sweeps (a lazily updated running
    # maximum in the flash kernel, none at all in the single-pass stream kernel that takes the windowed call when the
    # q/k norms allow it): the two results differ by that decorrelated rounding noise only -- both sit at the same
    # distance from fp32 torch math, and the fp32 log-sum-exp agrees to rounding
    assert _rel(fast, full) < 4e-3, _rel(fast, full)
    assert _rel(lse_fast, lse_full) < 1e-5
    ref = _attn_ref(qkv, batch, seq, heads, None, slopes, scale)
    e_fast, e_full = _rel(fast, ref), _rel(full, ref)
    # the flash kernel rounds P relative to (nearly) the row maximum, so a row's dominant probabilities are 1.0 or close
    # to it and round exactly; the stream kernel's P carry a plain bf16 rounding error (2.3e-3 at every slope)
    assert e_fast < 3e-3 and e_fast < 1.3 * e_full + 1e-4, (e_fast, e_full)


def test_attention_stream_and_flash_kernels_split_the_heads_by_norm():
    """Windowed forward on contiguous sequences: the single-pass stream kernel (no running maximum) takes the heads whose
    max|q| max|k| allows a fixed exponent reference, the flash kernel (second launch, head filter) the others. Heads of
    both kinds in one call, with dropout, a ragged length, and a zero-slope head; every head against fp32 torch math."""
    from animal2vec_b200 import ops

    batch, seq, heads = 2, 1111, 8
    d = heads * 64
    qkv = torch.randn(batch, seq, 3 * d, device="cuda", generator=_g(1))
    big = torch.tensor([1.0, 6.0, 1.0, 0.3, 8.0, 1.0, 5.0, 1.0], device="cuda")  # heads 1, 4, 6: 2 B >> 96 log2 units
    qkv.view(batch, seq, 3, heads, 64)[:, :, :2] *= big.view(1, 1, 1, heads, 1)
    qkv = qkv.bfloat16()
    slopes = torch.tensor([2.0 ** (-8.0 * (h + 1) / heads) for h in range(heads)], device="cuda")
    scale = torch.rand(heads, device="cuda", generator=_g(3)) + 0.5
    scale[2] = 0.0
    out, lse = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale, skip_far_keys=True)
    ref = _attn_ref(qkv, batch, seq, heads, None, slopes, scale)
    e = ((out.float() - ref).view(batch, seq, heads, 64).pow(2).sum((0, 1, 3)) /
         ref.view(batch, seq, heads, 64).pow(2).sum((0, 1, 3))).sqrt()
    assert float(e.max()) < 1e-2, e
    q, k, _ = qkv.float().view(batch, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    pos = torch.arange(seq, device="cuda").float()
    s = (q * 64 ** -0.5) @ k.transpose(-1, -2) - (slopes * scale).view(1, heads, 1, 1) * (pos[:, None] - pos[None, :]).abs()
    assert _rel(lse, torch.logsumexp(s, -1)) < 1e-4
    # dropout: both kernels regenerate the same keep hashes, so the windowed call equals the full sweep of the flash kernel
    a, _ = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale, drop_p=0.2, seed=77, skip_far_keys=True)
    b, _ = ops.attn_fwd(qkv, batch, seq, heads, slopes=slopes, alibi_scale=scale, drop_p=0.2, seed=77)
    assert _rel(a, b) < 8e-3, _rel(a, b)


def test_attention_dropout_statistics():
    from animal2vec_b200 import ops

    batch, seq, heads = 2, 142, 2
    d = heads * 64
    qkv = torch.zeros(batch, seq, 3 * d, device="cuda", dtype=torch.bfloat16)
    qkv[:, :, 2 * d:] = 1.0  # v = 1, uniform attention -> out = fraction kept / (1-p)
    out, _ = ops.attn_fwd(qkv, batch, seq, heads, drop_p=0.1, seed=7)
    assert abs(out.float().mean().item() - 1.0) < 0.01
    assert out.float().std().item() > 1e-3
    out2, _ = ops.attn_fwd(qkv, batch, seq, heads, drop_p=0.1, seed=7)
    assert torch.equal(out, out2)


# --------------------------------------------------------------------------------------- masking
def test_mask_index_and_gathers():
    from animal2vec_b200 import ops

    b, m, t, d, tk = 3, 4, 200, 64, 37
    rows = b * m
    mask = torch.ones(rows, t, dtype=torch.uint8)
    g = torch.Generator().manual_seed(0)
    for r in range(rows):
        mask[r, torch.randperm(t, generator=g)[:tk]] = 0
    mask = mask.cuda()
    mi = ops.mask_index(mask, tk, m)
    assert mi.err.item() == 0
    keep_ref = torch.stack([torch.nonzero(mask[r] == 0).flatten() for r in range(rows)])
    assert torch.equal(mi.ids_keep.long(), keep_ref)
    # ids_restore: gather of [kept..., masked...] by ids_restore restores the order
    cat = torch.cat([keep_ref, torch.stack([torch.nonzero(mask[r]).flatten() for r in range(rows)])], 1)
    assert torch.equal(torch.gather(cat, 1, mi.ids_restore.long()), torch.arange(t, device="cuda").expand(rows, t))

    x = torch.randn(b, t, d, device="cuda", generator=_g(1))
    xc = x.repeat_interleave(m, 0)
    x_masked = ops.row_gather(x.view(-1, d), mi.clone_src, rows * t, out_shape=(rows, t, d))
    assert torch.equal(x_masked, xc * (1 - mask.float()).unsqueeze(-1))
    x_unm = ops.row_gather(x.view(-1, d), mi.keep_src_x, rows * tk, out_shape=(rows, tk, d))
    assert torch.equal(x_unm, torch.gather(xc, 1, keep_ref.unsqueeze(-1).expand(-1, -1, d)))
    pos = torch.randn(rows, t, d, device="cuda", generator=_g(2))
    added = ops.row_gather(pos.view(-1, d), mi.keep_src_clone, rows * tk, add=x_unm, out_shape=(rows, tk, d))
    assert torch.allclose(added, x_unm + torch.gather(pos, 1, keep_ref.unsqueeze(-1).expand(-1, -1, d)))
    # decoder input: scatter kept rows back, zeros (std 0) elsewhere
    dec = ops.row_gather(x_unm.view(-1, d), mi.restore_src, rows * t, out_shape=(rows, t, d))
    ref = torch.zeros(rows, t, d, device="cuda")
    ref.scatter_(1, keep_ref.unsqueeze(-1).expand(-1, -1, d), x_unm)
    assert torch.equal(dec, ref)
    # noise fill statistics
    dec_n = ops.row_gather(x_unm.view(-1, d), mi.restore_src, rows * t, out_shape=(rows, t, d), fill_std=0.01,
                           fill_seed=5)
    noise = dec_n[mask.bool()]
    assert abs(noise.std().item() - 0.01) < 5e-4 and abs(noise.mean().item()) < 2e-4
    assert torch.equal(dec_n[~mask.bool()], ref[~mask.bool()])
    # clone-sum backward
    d_m = torch.randn(rows, t, d, device="cuda", generator=_g(3))
    d_u = torch.randn(rows, tk, d, device="cuda", generator=_g(4))
    dx = ops.clone_sum_bwd(d_m, d_u, mi.restore_src, b, t, m, d)
    ref_dx = (d_m * (1 - mask.float()).unsqueeze(-1))
    ref_dx = ref_dx + torch.zeros_like(d_m).scatter_(1, keep_ref.unsqueeze(-1).expand(-1, -1, d), d_u)
    ref_dx = ref_dx.view(b, m, t, d).sum(1)
    assert _rel(dx, ref_dx) < 1e-6
    # bad Tk is flagged
    assert ops.mask_index(mask, tk + 1, m).err.item() != 0


# --------------------------------------------------------------------------------------- targets / loss
@pytest.mark.parametrize("b,m,t,d", [(2, 12, 333, 1024), (3, 5, 100, 256), (1, 7, 2000, 512)])
def test_loss_fused_forward_backward_in_place(b, m, t, d):
    """a2v_d2v_loss_fused (bf16): loss, the four column statistics and the in-place gradient against torch and against
    the separate forward / backward kernels; wide target kernels against F.instance_norm."""
    from animal2vec_b200 import ops

    r = b * m
    pred = torch.randn(r, t, d, device="cuda", generator=_g(1)).bfloat16()
    y = torch.randn(b, t, d, device="cuda", generator=_g(2))
    mask = (torch.rand(r, t, device="cuda", generator=_g(3)) < 0.9).to(torch.uint8)
    scale = d ** -0.5
    l0, st0 = ops.d2v_loss_fwd(pred, y, mask, m, scale)  # dispatches to the fused kernel without gradient
    g0 = ops.d2v_loss_bwd(pred, y, mask, m, scale, None)
    mb = mask.bool()
    yc = y.repeat_interleave(m, 0)
    xs, ys = pred.float()[mb], yc[mb]
    ref_loss = ((xs - ys) ** 2).sum().double() * scale
    assert abs(float(l0) - float(ref_loss)) <= 1e-5 * float(ref_loss)
    for k, v in enumerate((xs.sum(0), (xs * xs).sum(0), ys.sum(0), (ys * ys).sum(0))):
        assert _rel(st0[k], v) < 1e-5, k
    work = pred.clone()
    l1, st1 = ops.d2v_loss_fused(work, y, mask, m, scale, 2.0 * scale)
    assert abs(float(l1) - float(l0)) <= 1e-9 * abs(float(l0)) + 1e-9 and _rel(st1, st0) < 1e-7
    ref_g = (2 * scale * (pred.float() - yc)) * mb.unsqueeze(-1)
    assert _rel(work, ref_g) < 4e-3 and _rel(work, g0) < 1e-6
    assert float(work[~mb].abs().sum()) == 0.0



@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_targets_and_loss(dtype):
    from animal2vec_b200 import ops

    k, b, t, d, m = 5, 2, 333, 256, 3
    layers = [(torch.randn(b, t, d, device="cuda", generator=_g(i)) * (1 + i) + i).to(dtype) for i in range(k)]
    y = ops.make_targets(layers)
    ref = sum(F.instance_norm(x.float().transpose(1, 2)).transpose(1, 2) for x in layers) / k
    assert _rel(y, ref) < (1e-5 if dtype == torch.float32 else 1e-5), _rel(y, ref)

    rows = b * m
    pred = torch.randn(rows, t, d, device="cuda", generator=_g(20)).to(dtype)
    mask = (torch.rand(rows, t, device="cuda", generator=_g(21)) < 0.9).to(torch.uint8)
    scale = 1 / math.sqrt(d)
    loss, stats = ops.d2v_loss_fwd(pred, y, mask, m, scale)
    yb = y.repeat_interleave(m, 0)[mask.bool()]
    xb = pred.float()[mask.bool()]
    ref_loss = (F.mse_loss(xb, yb, reduction="none") * scale).sum()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-5
    n = xb.shape[0]
    sx, sxx, sy, syy = stats
    var_x = (sxx - sx * sx / n) / (n - 1)
    assert _rel(torch.sqrt(var_x + 1e-6).mean(), torch.sqrt(xb.double().var(dim=0) + 1e-6).mean()) < 1e-5
    var_y = (syy - sy * sy / n) / (n - 1)
    assert _rel(torch.sqrt(var_y + 1e-6).mean(), torch.sqrt(yb.double().var(dim=0) + 1e-6).mean()) < 1e-5
    g = torch.tensor([0.37], device="cuda")
    dpred = ops.d2v_loss_bwd(pred, y, mask, m, scale, g)
    ref_d = torch.zeros(rows, t, d, device="cuda")
    ref_d[mask.bool()] = 2 * scale * 0.37 * (xb - yb)
    assert _rel(dpred, ref_d) < (1e-6 if dtype == torch.float32 else 4e-3)


# --------------------------------------------------------------------------------------- utilities
def test_util_kernels():
    from animal2vec_b200 import gemm, ops

    x = torch.randn(1001, 384, device="cuda", generator=_g(1))
    out = torch.ones(384, device="cuda")
    ops.colsum(x, out)
    assert _rel(out, 1 + x.sum(0)) < 1e-5
    ops.colsum(x.bfloat16(), out.zero_())
    assert _rel(out, x.bfloat16().float().sum(0)) < 1e-5
    for rows in (1, 37, 5003):  # bf16, C % 256 == 0: the 16-byte / four-rows-in-flight kernel
        xw = torch.randn(rows, 768, device="cuda", generator=_g(3)).bfloat16()
        ow = torch.ones(768, device="cuda")
        ops.colsum(xw, ow)
        assert _rel(ow, 1 + xw.float().sum(0)) < 1e-5

    w = torch.randn(6, 5, 7, device="cuda", generator=_g(2))  # (O, I, k)
    w_fwd = ops.cast_strided(w, (6, 7, 5), (35, 1, 7), out_dtype=torch.float32)
    assert torch.equal(w_fwd.view(6, 7, 5), w.permute(0, 2, 1))
    w_flip = ops.cast_strided(w, (5, 7, 6), (7, -1, 35), offset=6, out_dtype=torch.bfloat16)
    assert torch.equal(w_flip.view(5, 7, 6), w.flip(2).permute(1, 2, 0).bfloat16())

    # hi/lo split GEMM reaches fp32-class accuracy
    a = torch.randn(300, 256, device="cuda", generator=_g(3))
    b = torch.randn(192, 256, device="cuda", generator=_g(4))
    c = gemm.gemm_nt(ops.split3(a, 0), ops.split3(b, 1), out_dtype=torch.float32)
    assert _rel(c, a.double() @ b.double().t()) < 3e-5
    ct = torch.zeros(256, 256, device="cuda")
    a2 = torch.randn(300, 256, device="cuda", generator=_g(5))
    gemm.gemm_tn(ops.split3(a, 2), ops.split3(a2, 3), ct)
    assert _rel(ct, a.double().t() @ a2.double()) < 3e-5

    n = 4096 * 3 + 4
    s, e = torch.randn(n, device="cuda", generator=_g(6)), torch.randn(n, device="cuda", generator=_g(7))
    e0 = e.clone()
    lp = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ops.ema_step(s, e, lp, 0.9997)
    ref = e0 * 0.9997 + s * (1 - 0.9997)
    assert torch.allclose(e, ref, rtol=1e-6, atol=1e-7) and torch.equal(lp, e.bfloat16())

    p, g_ = torch.randn(n, device="cuda", generator=_g(8)), torch.randn(n, device="cuda", generator=_g(9))
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pr = p.clone()
    plp = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    gs = torch.tensor([0.5], device="cuda")
    mr, vr = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in (1, 2, 3):
        ops.adamw_step(p, g_, m, v, plp, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.01, step=step,
                       grad_scale=gs)
        gg = g_ * 0.5
        mr = 0.9 * mr + 0.1 * gg
        vr = 0.98 * vr + 0.02 * gg * gg
        pr = pr - 1e-3 * 0.01 * pr
        pr = pr - (1e-3 * math.sqrt(1 - 0.98 ** step) / (1 - 0.9 ** step)) * mr / (vr.sqrt() + 1e-6)
    assert torch.allclose(p, pr, rtol=1e-5, atol=1e-6) and torch.equal(plp, p.bfloat16())

    ss = torch.zeros(1, device="cuda", dtype=torch.float64)
    ops.sumsq(g_, ss)
    assert abs(ss.item() - (g_.double() ** 2).sum().item()) / ss.item() < 1e-6
    out2 = torch.zeros(2, device="cuda")
    denom = torch.tensor([250.0], device="cuda")
    ops.clip_coef(ss, denom, 2.0, 1.0, out2)
    norm = 2.0 / 250.0 * math.sqrt(ss.item())
    assert abs(out2[1].item() - norm) / norm < 1e-5
    assert abs(out2[0].item() - 2.0 / 250.0 * min(1.0, 1.0 / (norm + 1e-6))) < 1e-7


# --------------------------------------------------------------------------------------- sinc / im2col / mixup
def _sinc_buffers(k=63, sr=8000):
    n_lin = torch.linspace(0, (k / 2) - 1, steps=int(k / 2))
    window = 0.53836 - 0.46164 * torch.cos(2 * math.pi * n_lin / k)
    n = (k - 1) / 2.0
    n_ = 2 * math.pi * torch.arange(-n, 0).view(1, -1) / sr
    return n_.float().cuda(), window.float().cuda()


def _sinc_filters_ref(low_hz, band_hz, n_, window, k, min_low, min_band, sr):
    low = min_low + low_hz.abs()
    high = torch.clamp(low + min_band + band_hz.abs(), min_low, sr / 2)
    band = (high - low)[:, 0]
    left = (torch.sin(high @ n_) - torch.sin(low @ n_)) / n_ * 2 * window
    bp = torch.cat([left, 2 * band.view(-1, 1), left.flip(1)], 1)
    return bp / (2 * band[:, None])


def test_sinc():
    from animal2vec_b200 import ops

    c, k, sr, b, n = 127, 63, 8000, 2, 3000
    n_, window = _sinc_buffers(k, sr)
    mel = torch.linspace(2595 * math.log10(1 + 50 / 700), 2595 * math.log10(1 + (4000 - 50 - 127) / 700), c + 1)
    hz = 700 * (10 ** (mel / 2595) - 1)
    low = hz[:-1].unsqueeze(1).cuda().requires_grad_(True)
    band = (hz[1:] - hz[:-1]).unsqueeze(1).cuda().requires_grad_(True)
    x = torch.randn(b, n, device="cuda", generator=_g(1))
    filt_ref = _sinc_filters_ref(low, band, n_, window, k, 50.0, 127.0, sr)
    y_ref = F.conv1d(F.pad(x.unsqueeze(1), (31, 31), mode="reflect"), filt_ref.view(c, 1, k))  # (B, C, N)
    dy = torch.randn(b, n, 128, device="cuda", generator=_g(2))
    dy[..., 127] = 0
    y_ref.backward(dy[..., :127].transpose(1, 2))

    filt = ops.sinc_filters_fwd(low.detach().view(-1), band.detach().view(-1), n_.view(-1), window, k, 50.0, 127.0, sr)
    assert _rel(filt[:127], filt_ref) < 1e-5 and filt[127].abs().sum().item() == 0
    y = ops.sinc_conv_fwd(x, filt, torch.float32)
    assert _rel(y[..., :127], y_ref.transpose(1, 2)) < 1e-5 and y[..., 127].abs().sum().item() == 0
    dfilt = ops.sinc_conv_wgrad(x, dy, k)
    filt_leaf = filt_ref.detach().clone().requires_grad_(True)
    F.conv1d(F.pad(x.unsqueeze(1), (31, 31), mode="reflect"), filt_leaf.view(c, 1, k)).backward(
        dy[..., :127].transpose(1, 2))
    assert _rel(dfilt[:127], filt_leaf.grad) < 1e-4
    dlow, dband = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    ops.sinc_filters_bwd(low.detach().view(-1), band.detach().view(-1), n_.view(-1), window, k, 50.0, 127.0, sr, dfilt,
                         dlow, dband)
    assert _rel(dlow, low.grad.view(-1)) < 2e-3, _rel(dlow, low.grad.view(-1))
    assert _rel(dband, band.grad.view(-1)) < 2e-3, _rel(dband, band.grad.view(-1))


@pytest.mark.parametrize("k,n", [(63, 3000), (125, 1100), (31, 257)])
def test_sinc_conv_forward_on_tensor_cores(k, n):
    """bf16 output path: Toeplitz operand x filter matrix on tcgen05 with hi + lo bf16 splits of both operands (three
    products); against fp32 conv1d with reflect padding, ragged lengths, tap counts below / above one 64-tap half."""
    from animal2vec_b200 import ops

    b, c = 3, 127
    x = torch.randn(b, n, device="cuda", generator=_g(1)) * 3.0
    filt = torch.zeros(128, k, device="cuda")
    filt[:c] = torch.randn(c, k, device="cuda", generator=_g(2)) * torch.hann_window(k, device="cuda")
    ref = F.conv1d(F.pad(x.unsqueeze(1), (k // 2, k // 2), mode="reflect"), filt[:c].view(c, 1, k)).transpose(1, 2)
    y16 = ops.sinc_conv_fwd(x, filt, torch.bfloat16)
    y32 = ops.sinc_conv_fwd(x, filt, torch.float32)
    assert _rel(y32[..., :c], ref) < 1e-5
    assert y16[..., 127].float().abs().sum().item() == 0
    # the split products carry ~2^-17 relative operand error: the result differs from the fp32 kernel's by far less than
    # the bf16 rounding of the output
    assert _rel(y16[..., :c], ref) < 3e-3, _rel(y16[..., :c], ref)
    assert _rel(y16.float(), y32.bfloat16().float()) < 1e-3, _rel(y16.float(), y32.bfloat16().float())
    # filter gradient with bf16 dy: dy^T x Toeplitz operand on tcgen05 (x split hi + lo), accumulated on top of nothing
    dy = torch.randn(b, n, 128, device="cuda", generator=_g(3)).bfloat16()
    dy[..., 127] = 0
    leaf = filt[:c].clone().requires_grad_(True)
    F.conv1d(F.pad(x.unsqueeze(1), (k // 2, k // 2), mode="reflect"), leaf.view(c, 1, k)).backward(
        dy[..., :c].float().transpose(1, 2))
    dfilt = ops.sinc_conv_wgrad(x, dy, k)
    assert _rel(dfilt[:c], leaf.grad) < 1e-4, _rel(dfilt[:c], leaf.grad)
    assert dfilt[127].abs().sum().item() == 0


@pytest.mark.parametrize("k,s,pad,c,tin", [(10, 5, 3, 128, 1000), (3, 2, 1, 512, 401), (3, 2, 1, 64, 400)])
def test_im2col_col2im(k, s, pad, c, tin):
    from animal2vec_b200 import ops

    b = 2
    tout = (tin + 2 * pad - k) // s + 1
    x = torch.randn(b, tin, c, device="cuda", generator=_g(1))
    col = ops.im2col(x, k, s, pad, tout)
    w = torch.randn(32, c, k, device="cuda", generator=_g(2))
    ref = F.conv1d(x.transpose(1, 2), w, stride=s, padding=pad).transpose(1, 2)
    got = col.view(b * tout, k * c) @ w.permute(0, 2, 1).reshape(32, k * c).t()
    assert _rel(got.view(b, tout, 32), ref) < 1e-5
    dcol = torch.randn(b, tout, k * c, device="cuda", generator=_g(3))
    dx = ops.col2im(dcol, k, s, pad, tin)
    xr = x.clone().requires_grad_(True)
    unf = F.unfold(F.pad(xr.transpose(1, 2), (pad, pad)).unsqueeze(-1), (k, 1), stride=(s, 1))  # (B, C*k, Tout)
    unf = unf.view(b, c, k, -1).permute(0, 3, 2, 1).reshape(b, -1, k * c)[:, :tout]
    unf.backward(dcol)
    assert _rel(dx, xr.grad) < 1e-6


def test_mixup():
    from animal2vec_b200 import ops
    import numpy as np

    b, n, fs, wl = 3, 8000, 8000, 0.05
    n_fft = round(fs * wl)
    x = torch.randn(b, n, device="cuda", generator=_g(1)) * torch.tensor([1.0, 0.3, 2.0], device="cuda").view(-1, 1)
    freq = np.linspace(0, fs // 2, n_fft // 2 + 1)
    fsq = freq ** 2
    fsq[0] = 1.0
    wdb = 2.0 + 20.0 * (2 * np.log10(12194) + 2 * np.log10(fsq) - np.log10(fsq + 12194 ** 2) - np.log10(fsq + 20.6 ** 2)
                        - 0.5 * np.log10(fsq + 107.7 ** 2) - 0.5 * np.log10(fsq + 737.9 ** 2))
    aw = torch.from_numpy(np.power(10, np.maximum(wdb, -80.0) / 10)).cuda()
    hann = torch.hann_window(n_fft, device="cuda")
    fr = x.unfold(-1, n_fft, n_fft // 2)
    g_ref = ((torch.fft.rfft(hann * fr).abs() ** 2) * aw).sum(-1)
    gdb_ref = 10 * torch.log10(torch.maximum(g_ref, torch.tensor(10 ** (-80.0 / 10), device="cuda")))
    gdb = ops.mixup_gain(x, hann, aw.float(), n_fft, n_fft // 2)
    assert (gdb - gdb_ref.float()).abs().max().item() < 1e-3
    perm = torch.tensor([2, 0, 1], device="cuda", dtype=torch.int32)
    r = 0.7
    out, p = ops.mixup_apply(x, perm, gdb, r)
    G1 = gdb_ref.max(-1).values.float()
    G2 = G1[perm.long()]
    p_ref = (1 / (1 + 10 ** ((G1 - G2) / 20) * (1 - r) / r)).unsqueeze(-1)
    ref = (p_ref * x + (1 - p_ref) * x[perm.long()]) / torch.sqrt(p_ref ** 2 + (1 - p_ref) ** 2)
    assert _rel(out, ref) < 1e-4


# --------------------------------------------------------------------------------------- batched re-layout, fused dgelu
def test_relayout_batch_matches_single_launches():
    """One table-driven launch == the per-item launches: Linear transposes (tiled path), tap-major conv packs
    (generic path, padded), and gradient unpacking with accumulate + source clearing."""
    from animal2vec_b200 import ops, params as P

    dev = "cuda"
    packs = [P.pack_linear_t("a", 96, 160), P.pack_linear_t("b", 1024, 512), P.pack_conv_fwd("c", 4, 48, 64, 7, ngp=64, cgp=64),
             P.pack_conv_dgrad("d", 4, 64, 64, 19), P.pack_cols_padded_t("e", 128, 4, 48, 64), P.pack_bias_padded("f", 4, 48, 64)]
    srcs, ref, table, outs = [], [], ops.RelayoutTable(dev), []
    for i, pk in enumerate(packs):
        n_src = sum(max(0, (d_ - 1) * s_) for d_, s_ in zip(pk.dims, pk.in_strides)) + 1 + pk.in_off
        src = torch.randn(n_src, device=dev, generator=_g(10 + i))
        dt = torch.float32 if pk.is_bias else torch.bfloat16
        a = torch.zeros(pk.out_shape, device=dev, dtype=dt)
        b = torch.zeros(pk.out_shape, device=dev, dtype=dt)
        ops.relayout(src, a, pk.dims, pk.in_strides, pk.in_off, pk.out_strides, 0)
        table.add(src, b, pk.dims, pk.in_strides, pk.in_off, pk.out_strides, 0)
        srcs.append(src); ref.append(a); outs.append(b)
    table.run()
    for a, b in zip(ref, outs):
        assert torch.equal(a, b)
    # inverse direction: G += packed (fp32), packed cleared
    pk = packs[2]
    packed = torch.randn(pk.gt_shape, device=dev, generator=_g(30))
    g0 = torch.randn(4 * 48 * 64 * 7, device=dev, generator=_g(31))
    want = g0.clone()
    ops.relayout(packed, want, pk.dims, pk.gt_strides, 0, pk.in_strides, pk.in_off, accumulate=True)
    got = g0.clone()
    t2 = ops.RelayoutTable(dev)
    p2 = packed.clone()
    t2.add(p2, got, pk.dims, pk.gt_strides, 0, pk.in_strides, pk.in_off, accumulate=True, zero_src=True)
    t2.run()
    assert torch.equal(want, got)
    # every REAL source element was cleared
    probe = torch.zeros_like(want)
    ops.relayout(p2, probe, pk.dims, pk.gt_strides, 0, pk.in_strides, pk.in_off)
    assert probe.abs().sum().item() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dgelu_mul_fused_colsum(dtype):
    from animal2vec_b200 import ops

    rows, c = 3001, 4096
    dh = torch.randn(rows, c, device="cuda", generator=_g(1)).to(dtype)
    u = torch.randn(rows, c, device="cuda", generator=_g(2)).to(dtype)
    ur = u.float().clone().requires_grad_(True)
    F.gelu(ur).backward(dh.float())
    want = ops.dgelu_mul(dh.clone(), u)
    cs = torch.zeros(c, device="cuda")
    got = ops.dgelu_mul(dh.clone(), u, colsum=cs)
    assert torch.equal(want, got)
    assert _rel(got.float(), ur.grad) < (1e-5 if dtype == torch.float32 else 8e-3)
    assert _rel(cs, got.float().sum(0)) < 1e-5
