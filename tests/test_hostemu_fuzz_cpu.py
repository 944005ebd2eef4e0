"""Randomised shapes through the C ABI on the host-executed CUDA sources (tests/hostemu): the tensor-core convolution forward /
weight gradient on the functional tcgen05 / TMA model and a set of HBM-bound kernels, each against torch fp32 on the same 16-bit
operands.  The fixed GPU test lists cover the shapes the models use; this sweeps what they do not — ragged widths and heights,
single rows, odd source splits, channel slices of wider buffers, output channel counts that are not a multiple of the tile — where a
descriptor, clipping or indexing slip would hide.  An illegal shape must be REJECTED with ValueError / NotImplementedError (the
wrapper's argument checks), never mis-computed: every accepted call is checked.

Seeds are fixed; GDL_HOSTEMU_FUZZ=<n> sets the number of cases per family (default 12).
"""
import os
import random

import pytest
import torch
import torch.nn.functional as F

import hostemu

CASES = int(os.environ.get("GDL_HOSTEMU_FUZZ", "12"))


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


def _sources(rng, g, n, h, w, dt):
    """1-3 sources; each either a dense tensor or a channel slice of a wider buffer (ld > C, 16-byte aligned offset)"""
    srcs = []
    for _ in range(rng.choice([1, 1, 2, 3])):
        c = rng.choice([16, 16, 32, 48, 64, 80, 96, 128, 160, 256])  # the conv kernels take sources in multiples of 16 channels
        if rng.random() < 0.4:
            wide = (torch.randn(n, h, w, c + 24, generator=g) * 0.5).to(dt)
            srcs.append(wide[..., 8:8 + c])
        else:
            srcs.append((torch.randn(n, h, w, c, generator=g) * 0.5).to(dt))
    return srcs


@pytest.mark.parametrize("seed", range(CASES))
def test_conv_forward_random_shapes(monkeypatch, seed):
    hostemu.install(monkeypatch, torch_convs=False)
    from gdl_b200 import ops
    rng = random.Random(1000 + seed)
    g = torch.Generator().manual_seed(seed)
    dt = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    r = rng.choice([1, 1, 3, 3, 3, 5, 7])
    pad = rng.choice([0, r // 2])
    n, h, w = rng.choice([1, 2, 3]), rng.randint(max(1, r - 2 * pad), 20), rng.randint(max(1, r - 2 * pad), 150)
    srcs = _sources(rng, g, n, h, w, dt)
    ctot = sum(t.shape[3] for t in srcs)
    cout = rng.choice([1, 4, 5, 8, 16, 19, 32, 40, 64, 96, 128, 136, 256])
    wt = torch.randn(cout, ctot, r, r, generator=g) / (ctot * r * r) ** 0.5
    bias = torch.randn(cout, generator=g) if rng.random() < 0.5 else None
    relu = rng.random() < 0.5
    out_dtype = rng.choice([dt, dt, torch.float32])
    wp = ops.pack_conv_weight(wt, dt)
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    w32 = wp.view(cout, r, r, ctot).float().permute(0, 3, 1, 2)
    ref = F.conv2d(x, w32, bias, padding=pad)
    ho, wo = ref.shape[2], ref.shape[3]
    residual = None
    if rng.random() < 0.3 and cout % 8 == 0:
        residual = torch.randn(n, ho, wo, cout, generator=g).to(rng.choice([dt, torch.float32]))
        ref = ref + residual.float().permute(0, 3, 1, 2)
    if relu:
        ref = F.relu(ref)
    # sometimes write into a channel slice of a wider output buffer
    out = None
    if rng.random() < 0.3 and cout % 8 == 0:
        wide = torch.full((n, ho, wo, cout + 16), 7.0).to(out_dtype)
        out = wide[..., 8:8 + cout]
    try:
        y = ops.conv2d_fwd(srcs, wp, cout, r, r, pad, pad, out=out, out_dtype=out_dtype, bias=bias, relu=relu, residual=residual)
    except (ValueError, NotImplementedError) as e:
        pytest.skip(f"rejected by the wrapper: {e}")
    tol = 2e-3 if out_dtype == torch.float32 else (2.0 ** -7 if out_dtype == torch.bfloat16 else 2.0 ** -9)
    assert _rel(y, ref.permute(0, 2, 3, 1)) < tol, (n, h, w, [t.shape[3] for t in srcs], cout, r, pad, dt, out_dtype)
    if out is not None:  # neighbours of the slice untouched
        assert (wide[..., :8].float() == 7.0).all() and (wide[..., 8 + cout:].float() == 7.0).all()


@pytest.mark.parametrize("seed", range(CASES))
def test_conv_wgrad_random_shapes(monkeypatch, seed):
    hostemu.install(monkeypatch, torch_convs=False)
    from gdl_b200 import ops
    rng = random.Random(2000 + seed)
    g = torch.Generator().manual_seed(100 + seed)
    dt = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    r = rng.choice([1, 3, 3, 3])
    pad = r // 2
    n, h, w = rng.choice([1, 2]), rng.randint(1, 14), rng.randint(1, 140)
    srcs = _sources(rng, g, n, h, w, dt)
    ctot = sum(t.shape[3] for t in srcs)
    cout = rng.choice([8, 16, 32, 48, 64, 128, 160])
    dy = (torch.randn(n, h, w, cout, generator=g) * 0.5).to(dt)
    dw = torch.zeros(cout, r * r * ctot)
    try:
        ops.conv2d_wgrad(srcs, dy, r, r, pad, pad, dw)
    except (ValueError, NotImplementedError) as e:
        pytest.skip(f"rejected by the wrapper: {e}")
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(x, (cout, ctot, r, r), dy.float().permute(0, 3, 1, 2).contiguous(), padding=pad)
    assert _rel(dw.view(cout, r, r, ctot), ref.permute(0, 2, 3, 1)) < 2e-3, (n, h, w, [t.shape[3] for t in srcs], cout, r, dt)


@pytest.mark.parametrize("seed", range(CASES))
def test_hbm_kernels_random_shapes(monkeypatch, seed):
    hostemu.install(monkeypatch, torch_convs=False)
    from gdl_b200 import ops
    rng = random.Random(3000 + seed)
    g = torch.Generator().manual_seed(200 + seed)
    dt = rng.choice([torch.bfloat16, torch.float16])
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    n, h, w, c = rng.choice([1, 2, 3]), rng.randint(1, 17), rng.randint(1, 23), 8 * rng.randint(1, 12)
    x = torch.randn(n, h, w, c, generator=g).to(dt)
    # bilinear resize to an arbitrary size (align_corners=False) and its adjoint
    ho, wo = rng.randint(1, 40), rng.randint(1, 40)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=False)
    y = ops.bilinear_fwd(x, ho, wo)
    assert _rel(y, ref.permute(0, 2, 3, 1)) < 3 * ulp
    dy = torch.randn(n, ho, wo, c, generator=g).to(dt)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(ops.bilinear_bwd(dy, h, w), xr.grad.permute(0, 2, 3, 1)) < 4 * ulp
    # adaptive average pooling, any bin count
    s = rng.randint(1, 7)
    assert _rel(ops.adaptive_avgpool_fwd(x, s), F.adaptive_avg_pool2d(x.float().permute(0, 3, 1, 2), s).permute(0, 2, 3, 1)) < 2 * ulp
    # LayerNorm over the channel dim of token rows
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    yl, _ = ops.layernorm_fwd(x.view(-1, c), gamma, beta, 1e-6, dt, True)
    assert _rel(yl, F.layer_norm(x.float().view(-1, c), (c,), gamma, beta, 1e-6)) < 3 * ulp
    # attention-score softmax with key padding
    length = rng.randint(1, 70)
    lpad = (length + 15) // 16 * 16
    sc = (torch.randn(n * h, 2, lpad, generator=g) * 3).to(dt)
    p = ops.softmax_fwd(sc, 0.37, length)
    refp = (sc.float()[..., :length] * 0.37).softmax(-1)
    assert (p.float()[..., :length] - refp).abs().max() < 3 * ulp and not p[..., length:].any()
    # train-mode batch norm statistics + apply on a channel slice of a wider buffer
    wide = torch.randn(n, h, w, c + 8, generator=g).to(dt)
    xs = wide[..., 8:]
    sums = torch.zeros(2 * c)
    pivot = torch.randn(c, generator=g) * 0.1
    ops.bn_stats(xs, sums, pivot)
    d = xs.float().reshape(-1, c) - pivot
    assert torch.allclose(sums[:c], d.sum(0), rtol=1e-4, atol=1e-3) and torch.allclose(sums[c:], (d * d).sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("seed", range(CASES))
def test_fused_attention_random_shapes(monkeypatch, seed):
    """gdl_mha_flash_fwd / gdl_sra_attention_fwd on the functional model: token counts around the 128-query / 128-key block edges,
    few long CTAs (ring wrap, K / V reload), both dtypes"""
    hostemu.install(monkeypatch, torch_convs=False)
    from gdl_b200 import ops
    rng = random.Random(4000 + seed)
    g = torch.Generator().manual_seed(300 + seed)
    dt = rng.choice([torch.bfloat16, torch.float16])
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    b, heads = rng.choice([1, 2]), rng.choice([1, 2, 3])
    c = 64 * heads
    ops.set_option("sra_max_ctas", rng.choice([0, 0, 1, 2, 5]))
    try:
        if rng.random() < 0.6:
            n = rng.choice([1, 15, 16, 17, 127, 128, 129, 255, 256, 257, 383, 385, rng.randint(1, 700)])
            qkv = torch.randn(b * n, 3 * c, generator=g).to(dt)
            qkv[:, :c] *= rng.choice([1.0, 3.0])
            o = ops.mha_flash_fwd(qkv, b, n, heads, 0.125)
            q, k, v = (qkv[:, i * c:(i + 1) * c].float().view(b, n, heads, 64) for i in range(3))
            p = (torch.einsum("bnhd,bkhd->bhnk", q, k) * 0.125).softmax(-1)
            ref = torch.einsum("bhnk,bkhd->bnhd", p, v).reshape(b * n, c)
            assert (o.float() - ref).abs().max() <= 3 * ulp * ref.abs().max(), (b, n, heads, dt)
        else:
            n, nk = 128 * rng.randint(1, 4), 64 * rng.randint(1, 4)
            q = torch.randn(b, n, c, generator=g).to(dt)
            kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dt)
            save = rng.random() < 0.7
            o, p = ops.sra_attention_fwd(q, kv2, heads, nk, 0.125, save_p=save)
            k = kv2[:, :c].float().view(b, nk, heads, 64)
            v = kv2[:, c:].float().view(b, nk, heads, 64)
            pr = (torch.einsum("bnhd,bkhd->bhnk", q.float().view(b, n, heads, 64), k) * 0.125).softmax(-1)
            ref = torch.einsum("bhnk,bkhd->bnhd", pr, v).reshape(b, n, c)
            assert (o.float() - ref).abs().max() <= 3 * ulp * ref.abs().max(), (b, n, nk, heads, dt)
            if save:
                assert (p.float() - pr.permute(0, 2, 1, 3).reshape(b, n, heads * nk)).abs().max() <= 3 * ulp
    finally:
        ops.set_option("sra_max_ctas", 0)


@pytest.mark.parametrize("seed", range(max(4, CASES // 3)))
def test_fused_attention_backward_random_shapes(monkeypatch, seed):
    """gdl_sra_attention_bwd on the functional model (asynchronous completion, both operand-read modes): ragged query counts, 1-4 key
    chunks, few long CTAs"""
    hostemu.install(monkeypatch, torch_convs=False, async_seed=seed)
    from gdl_b200 import ops
    rng = random.Random(5000 + seed)
    g = torch.Generator().manual_seed(400 + seed)
    dt = rng.choice([torch.bfloat16, torch.float16])
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    b, heads = rng.choice([1, 2]), rng.choice([1, 2, 3])
    n, nk = rng.choice([64, 128, 200, 256, 300, 384]), 64 * rng.randint(1, 4)
    c = 64 * heads
    do = torch.randn(b, n, c, generator=g).to(dt)
    kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dt)
    p = (torch.randn(b, n, heads, nk, generator=g) * 2).softmax(-1).reshape(b, n, heads * nk).to(dt)
    ops.set_option("sra_max_ctas", rng.choice([0, 1, 2, 5]))
    try:
        dq, ds = ops.sra_attention_bwd(do, kv2, p, heads, nk, 0.125)
    finally:
        ops.set_option("sra_max_ctas", 0)
    k = kv2[:, :c].float().view(b, nk, heads, 64)
    v = kv2[:, c:].float().view(b, nk, heads, 64)
    pf = p.float().view(b, n, heads, nk)
    dp = torch.einsum("bnhd,bkhd->bnhk", do.float().view(b, n, heads, 64), v)
    ds_ref = 0.125 * pf * (dp - (pf * dp).sum(-1, keepdim=True))
    assert (ds.float().view(b, n, heads, nk) - ds_ref).abs().max() <= 3 * ulp * ds_ref.abs().max(), (b, n, nk, heads, dt)
    dq_ref = torch.einsum("bnhk,bkhd->bnhd", ds.float().view(b, n, heads, nk), k).reshape(b, n, c)
    assert (dq.float() - dq_ref).abs().max() <= 3 * ulp * dq_ref.abs().max(), (b, n, nk, heads, dt)
