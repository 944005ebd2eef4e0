"""Fused spatial-reduction attention forward (gdl_sra_attention_fwd, csrc/sra_attention.cu) against fp32 softmax(q.k^T).v on
the same 16-bit operands (reference: Attention.forward, mix_transformer.py:131-159) and against the three-kernel path it
replaces.  Tolerance: o and p are stored in 16 bits (one rounding, 2^-8 relative for bf16) and the second contraction
runs on 16-bit-rounded un-normalised probabilities: 3 ulp of the output scale is asserted.

Written after the round's GPU budget was spent: the kernel has been executed on the CPU functional model of the tcgen05 / TMA /
mbarrier features only (tests/test_hostemu_tensorcore_cpu.py runs this file) — its first run on a B200 is this test.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, kv2, heads, nk, scale):
    b, n, c = q.shape
    d = c // heads
    k = kv2[:, :c].float().view(b, nk, heads, d)
    v = kv2[:, c:].float().view(b, nk, heads, d)
    s = torch.einsum("bnhd,bkhd->bhnk", q.float().view(b, n, heads, d), k) * scale
    p = s.softmax(-1)
    o = torch.einsum("bhnk,bkhd->bnhd", p, v).reshape(b, n, c)
    return o, p.permute(0, 2, 1, 3).reshape(b, n, heads * nk)


@pytest.mark.parametrize("b,n,heads,nk,dtype,save_p", [
    (2, 256, 1, 256, torch.bfloat16, True),    # MiT stage 1 shape class (one head, 256 keys), two tiles per image
    (1, 128, 2, 64, torch.bfloat16, True),     # 256x256 tile: 64 keys
    (3, 128, 5, 128, torch.float16, True),     # 5 heads (stage 3), fp16, K/V reload between heads
    (1, 384, 2, 192, torch.bfloat16, False),   # inference: no probabilities written; 192 keys
    (2, 64, 8, 64, torch.bfloat16, True),      # MiT stage 4 of a 256x256 tile: 64 queries — half a query tile, rows clipped
    (1, 200, 2, 128, torch.bfloat16, True),    # ragged second query tile with saved probabilities
])
def test_fused_attention_matches_fp32(cuda, b, n, heads, nk, dtype, save_p):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(7)
    c = 64 * heads
    q = torch.randn(b, n, c, generator=g).to(dtype).cuda()
    kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dtype).cuda()
    scale = 64 ** -0.5
    assert ops.sra_attention_supported(n, nk, 64)
    o, p = ops.sra_attention_fwd(q, kv2, heads, nk, scale, save_p=save_p)
    torch.cuda.synchronize()
    o_ref, p_ref = _ref(q, kv2, heads, nk, scale)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert (o.float() - o_ref).abs().max() <= 3 * ulp * o_ref.abs().max()
    if save_p:
        assert (p.float() - p_ref).abs().max() <= 3 * ulp * p_ref.abs().max()
        assert (p.float().view(b, n, heads, nk).sum(-1) - 1).abs().max() < 4 * ulp * nk ** 0.5
    else:
        assert p is None


@pytest.mark.parametrize("ctas", [1, 3, 7])
def test_persistent_ctas_walk_tiles_heads_and_images(cuda, ctas):
    """few, long CTAs: Q double buffering wraps, K / V are re-loaded at head and image changes inside one CTA, barrier phases
    flip many times — the schedule of a full-size launch (2048 tiles over 148 CTAs) at test size"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(9)
    b, n, heads, nk, dt = 2, 384, 2, 64, torch.bfloat16
    c = 64 * heads
    q = torch.randn(b, n, c, generator=g).to(dt).cuda()
    kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dt).cuda()
    ops.set_option("sra_max_ctas", ctas)
    try:
        o, p = ops.sra_attention_fwd(q, kv2, heads, nk, 0.125)
        torch.cuda.synchronize()
    finally:
        ops.set_option("sra_max_ctas", 0)
    o_ref, p_ref = _ref(q, kv2, heads, nk, 0.125)
    assert (o.float() - o_ref).abs().max() <= 3 * 2.0 ** -8 * o_ref.abs().max()
    assert (p.float() - p_ref).abs().max() <= 3 * 2.0 ** -8 * p_ref.abs().max()


@pytest.mark.parametrize("b,n,heads,dtype,ctas", [
    (2, 197, 3, torch.bfloat16, 0),     # ViT-B/16 at 224: 196 patches + cls — ragged query tile and ragged second key block
    (1, 1297, 2, torch.bfloat16, 0),    # DOFA-base on a 512 tile: 36 x 36 patches + cls, 11 key blocks (the last holds 17 keys)
    (2, 300, 2, torch.float16, 2),      # two long CTAs: the K / V ring wraps across tiles, heads and images
    (1, 256, 1, torch.bfloat16, 0),     # exactly one resident block through the same entry point
    (1, 130, 1, torch.bfloat16, 0),     # 2 keys in the last block
])
def test_flash_self_attention_matches_fp32(cuda, b, n, heads, dtype, ctas):
    """gdl_mha_flash_fwd: self-attention of a fused qkv projection with the keys streamed in blocks of 128 (online softmax)"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(11)
    c = 64 * heads
    qkv = torch.randn(b * n, 3 * c, generator=g).to(dtype).cuda()
    qkv[:, :c] *= 2.0  # wider score range: the running maximum really moves between key blocks
    ops.set_option("sra_max_ctas", ctas)
    try:
        o = ops.mha_flash_fwd(qkv, b, n, heads, 0.125)
        torch.cuda.synchronize()
    finally:
        ops.set_option("sra_max_ctas", 0)
    q, k, v = (qkv[:, i * c:(i + 1) * c].float().view(b, n, heads, 64) for i in range(3))
    p = (torch.einsum("bnhd,bkhd->bhnk", q, k) * 0.125).softmax(-1)
    ref = torch.einsum("bhnk,bkhd->bnhd", p, v).reshape(b * n, c)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert (o.float() - ref).abs().max() <= 3 * ulp * ref.abs().max()


def test_dofa_encoder_with_flash_attention_equals_three_kernel_encoder(cuda):
    """the frozen-encoder route of the DOFA model (configs[3]) with the option on and off: same features up to 16-bit rounding"""
    from gdl_b200 import ops
    from gdl_b200.models.dofa import DOFAv2
    torch.manual_seed(0)
    enc = DOFAv2("dofa_base", img_size=112, depth=2, out_indices=[0, 1]).cuda().eval()
    with torch.no_grad():
        for name, prm in enc.named_parameters():
            if "ls1" in name or "ls2" in name:
                prm.fill_(0.5)  # LayerScale starts at 1e-5: give the attention branch weight in the features
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 112, 112, generator=g).cuda()
    wl = torch.tensor([0.665, 0.56, 0.49, 0.842]).cuda()
    feats = {}
    try:
        for flash in (0, 1):
            ops.set_option("mha_flash", flash)
            with torch.no_grad():
                feats[flash] = [f.float().clone() for f in enc(x, wl)]
    finally:
        ops.set_option("mha_flash", 0)
    torch.cuda.synchronize()
    for f0, f1 in zip(feats[0], feats[1]):
        assert ((f1 - f0).norm() / f0.norm()).item() < 1e-2


@pytest.mark.parametrize("b,n,heads,nk,dtype,ctas", [
    (2, 256, 1, 256, torch.bfloat16, 0),
    (1, 384, 2, 64, torch.bfloat16, 1),      # one long CTA: P tile reuse handshake, K / V reload between heads
    (2, 128, 5, 128, torch.float16, 3),
    (2, 200, 2, 64, torch.bfloat16, 2),      # ragged query tiles: zero-filled loads, clipped stores
])
def test_fused_attention_backward_matches_fp32(cuda, b, n, heads, nk, dtype, ctas):
    """gdl_sra_attention_bwd: dS = scale * P * (dP - rowsum(P * dP)), dQ = dS.K against torch fp32 on the same 16-bit operands and
    against the three launches it replaces"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(13)
    c = 64 * heads
    do = torch.randn(b, n, c, generator=g).to(dtype).cuda()
    kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dtype).cuda()
    p = (torch.randn(b, n, heads, nk, generator=g) * 2).softmax(-1).reshape(b, n, heads * nk).to(dtype).cuda()
    scale = 0.125
    ops.set_option("sra_max_ctas", ctas)
    try:
        dq, ds = ops.sra_attention_bwd(do, kv2, p, heads, nk, scale)
        torch.cuda.synchronize()
    finally:
        ops.set_option("sra_max_ctas", 0)
    k = kv2[:, :c].float().view(b, nk, heads, 64)
    v = kv2[:, c:].float().view(b, nk, heads, 64)
    pf = p.float().view(b, n, heads, nk)
    dp = torch.einsum("bnhd,bkhd->bnhk", do.float().view(b, n, heads, 64), v)
    ds_ref = scale * pf * (dp - (pf * dp).sum(-1, keepdim=True))
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert (ds.float().view(b, n, heads, nk) - ds_ref).abs().max() <= 3 * ulp * ds_ref.abs().max()
    dq_ref = torch.einsum("bnhk,bkhd->bnhd", ds.float().view(b, n, heads, nk), k).reshape(b, n, c)  # from the 16-bit dS the kernel used
    assert (dq.float() - dq_ref).abs().max() <= 3 * ulp * dq_ref.abs().max()
    # the three launches of models/segformer.py
    do4 = do.view(b, 1, n, c)
    dp3 = torch.empty((b, 1, n, heads * nk), dtype=dtype, device=do.device)
    ops.conv2d_fwd([do4[..., 0:64]], kv2[:, c:c + 64], nk, 1, 1, 0, 0, out=dp3[..., 0:nk], w_rows_per_img=nk, groups=(heads, 64, 64, nk))
    ds3 = ops.softmax_bwd(p.view(b, n, heads, nk), dp3.view(b, n, heads, nk), scale, nk)
    dq3 = torch.empty((b, 1, n, c), dtype=dtype, device=do.device)
    ops.conv2d_fwd([ds3.view(b, 1, n, heads * nk)[..., 0:nk]], kv2[:, 0:64], 64, 1, 1, 0, 0, out=dq3[..., 0:64], w_rows_per_img=nk,
                   w_mn_major=True, groups=(heads, nk, 64, 64))
    torch.cuda.synchronize()
    assert (ds.float() - ds3.view(b, n, heads * nk).float()).abs().max() <= 2.0 ** -5 * ds_ref.abs().max()
    assert (dq.float() - dq3.view(b, n, c).float()).abs().max() <= 2.0 ** -5 * dq_ref.abs().max()


def test_fused_attention_equals_three_kernel_path_and_rejects_other_shapes(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(8)
    b, n, heads, nk, dt = 2, 256, 2, 128, torch.bfloat16
    c = 64 * heads
    q = torch.randn(b, n, c, generator=g).to(dt).cuda()
    kv2 = torch.randn(b * nk, 2 * c, generator=g).to(dt).cuda()
    o, p = ops.sra_attention_fwd(q, kv2, heads, nk, 0.125)
    # the path it replaces (models/segformer.py): grouped q.k^T launch -> softmax -> grouped p.v launch
    q4 = q.view(b, 1, n, c)
    scores = torch.empty((b, 1, n, heads * nk), dtype=dt, device=q.device)
    ops.conv2d_fwd([q4[..., 0:64]], kv2[:, 0:64], nk, 1, 1, 0, 0, out=scores[..., 0:nk], w_rows_per_img=nk, groups=(heads, 64, 64, nk))
    p3 = ops.softmax_fwd(scores.view(b, n, heads, nk), 0.125, nk)
    o3 = torch.empty((b, 1, n, c), dtype=dt, device=q.device)
    ops.conv2d_fwd([p3.view(b, 1, n, heads * nk)[..., 0:nk]], kv2[:, c:c + 64], 64, 1, 1, 0, 0, out=o3[..., 0:64], w_rows_per_img=nk,
                   w_mn_major=True, groups=(heads, nk, 64, 64))
    torch.cuda.synchronize()
    # the three-kernel path rounds the scores to 16 bits before the softmax; the fused one keeps them in fp32
    assert (o.float() - o3.view(b, n, c).float()).abs().max() <= 2.0 ** -6 * o3.float().abs().max()
    assert (p.float() - p3.view(b, n, heads * nk).float()).abs().max() <= 2.0 ** -5 * p3.float().abs().max()
    assert not ops.sra_attention_supported(n, 144, 64) and not ops.sra_attention_supported(n, nk, 32)
    with pytest.raises(NotImplementedError):
        ops.sra_attention_fwd(q, kv2[: b * 48].contiguous(), heads, 48, 0.125)
    # without saved probabilities (inference) ragged query tiles and any key count are fine
    assert ops.sra_attention_supported(200, 48, 64, save_p=False)
    o_r, none = ops.sra_attention_fwd(q[:, :200].contiguous(), kv2[: b * 48].contiguous(), heads, 48, 0.125, save_p=False)
    torch.cuda.synchronize()
    o_ref, _ = _ref(q[:, :200], kv2[: b * 48], heads, 48, 0.125)
    assert none is None and (o_r.float() - o_ref).abs().max() <= 3 * 2.0 ** -8 * o_ref.abs().max()


def test_segformer_with_fused_attention_equals_three_kernel_model(cuda):
    """whole model (MiT-B1: head dim 64), 256x256 tile (64 keys per image at every stage): eval logits and a training step (loss +
    parameter gradients) with the option on and off"""
    from test_segformer_gpu import _rel, _setup
    import torch.nn.functional as F
    from gdl_b200 import ops
    prod = _setup("mit_b1", 3, 4)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3, 256, 256, generator=g).cuda()
    t = torch.randint(0, 4, (1, 256, 256), generator=g).cuda()
    state = {k: v.detach().clone() for k, v in prod.state_dict().items()}  # the train-mode forward moves BN running statistics
    calls = {"n": 0}
    real = ops.sra_attention_fwd

    def counted(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    ops.sra_attention_fwd = counted
    res = {}
    try:
        for fused in (0, 1):
            ops.set_option("sra_fused", fused)
            prod.load_state_dict(state)
            prod.eval()
            with torch.no_grad():
                ev = prod(x).float().clone()
            prod.train()
            prod.zero_grad(set_to_none=True)
            loss = F.cross_entropy(prod(x), t)
            loss.backward()
            res[fused] = (ev, loss.item(), {n: p.grad.detach().float().clone() for n, p in prod.named_parameters()})
    finally:
        ops.set_option("sra_fused", 0)
        ops.sra_attention_fwd = real
    torch.cuda.synchronize()
    # eval and train forward: all 4 stages x 2 blocks (stage 4's 64 queries are half a query tile); the backward of the 8 blocks
    # takes gdl_sra_attention_bwd
    assert calls["n"] == 8 + 8
    (e0, l0, g0), (e1, l1, g1) = res[0], res[1]
    # the two routes differ by 16-bit roundings of the scores / probabilities only
    flat = lambda gr: torch.cat([v.flatten() for v in gr.values()])  # noqa: E731
    print(f"fused vs three-kernel attention: logits {_rel(e1, e0):.4f}, loss {l0:.5f} / {l1:.5f}, all gradients {_rel(flat(g1), flat(g0)):.4f}")
    assert _rel(e1, e0) < 1e-2 and abs(l1 - l0) < 5e-3 * max(1.0, abs(l0))
    assert (e1.argmax(1) == e0.argmax(1)).float().mean() > 0.99
    assert _rel(flat(g1), flat(g0)) < 3e-2
