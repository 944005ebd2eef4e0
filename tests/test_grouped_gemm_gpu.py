"""GPU: grouped (all-heads-in-one-launch) attention GEMMs must equal the per-head launches bit for bit and match torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,n,nk,heads,d,dtype", [
    (3, 300, 256, 5, 64, torch.bfloat16),    # SegFormer stage 3 (sr-reduced keys)
    (2, 1024, 64, 2, 32, torch.float16),     # nk < one N tile
    (2, 200, 200, 3, 64, torch.bfloat16),    # DOFA-like: keys = queries, rows padded to 64
    (2, 77, 20, 8, 64, torch.bfloat16),      # nk not a multiple of 8: direct epilogue, masked columns
])
def test_grouped_attention_gemms(cuda, b, n, nk, heads, d, dtype):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(n + nk)
    c = heads * d
    lp = (nk + 63) // 64 * 64 if nk == n else (nk + 15) // 16 * 16
    q = (torch.randn(b, 1, n, c, generator=g) * 0.3).to(dtype).cuda()
    kv = (torch.randn(b * nk, 2 * c, generator=g) * 0.3).to(dtype).cuda()
    cout = lp if nk == n else nk  # the DOFA path writes whole padded rows
    s_loop = torch.zeros(b, 1, n, heads * lp, dtype=dtype, device="cuda")
    s_grp = torch.zeros_like(s_loop)
    for hd in range(heads):
        ops.conv2d_fwd([q[..., hd * d:(hd + 1) * d]], kv[:, hd * d:(hd + 1) * d], cout, 1, 1, 0, 0,
                       out=s_loop[..., hd * lp:hd * lp + cout], w_rows_per_img=nk)
    ops.conv2d_fwd([q[..., 0:d]], kv[:, 0:d], cout, 1, 1, 0, 0, out=s_grp[..., 0:cout], w_rows_per_img=nk,
                   groups=(heads, d, d, lp))
    assert torch.equal(s_loop, s_grp)
    kf = kv[:, :c].float().view(b, nk, heads, d)
    ref = torch.einsum("bnhd,bkhd->bnhk", q.float().view(b, n, heads, d), kf)
    got = s_grp.view(b, n, heads, lp)[..., :nk].float()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 6e-3

    p = ops.softmax_fwd(s_grp.view(b, n, heads, lp), d ** -0.5, nk).view(b, 1, n, heads * lp)
    o_loop = torch.zeros(b, 1, n, c, dtype=dtype, device="cuda")
    o_grp = torch.zeros_like(o_loop)
    for hd in range(heads):
        ops.conv2d_fwd([p[..., hd * lp:(hd + 1) * lp]], kv[:, c + hd * d:c + (hd + 1) * d], d, 1, 1, 0, 0,
                       out=o_loop[..., hd * d:(hd + 1) * d], w_rows_per_img=nk, w_mn_major=True)
    ops.conv2d_fwd([p[..., 0:lp]], kv[:, c:c + d], d, 1, 1, 0, 0, out=o_grp[..., 0:d], w_rows_per_img=nk,
                   w_mn_major=True, groups=(heads, lp, d, d))
    assert torch.equal(o_loop, o_grp)
    vf = kv[:, c:].float().view(b, nk, heads, d)
    ref_o = torch.einsum("bnhk,bkhd->bnhd", p.view(b, n, heads, lp)[..., :nk].float(), vf).reshape(b, 1, n, c)
    assert ((o_grp.float() - ref_o).abs().max() / ref_o.abs().max()).item() < 6e-3


@pytest.mark.parametrize("b,n,nk,heads,d,dtype", [
    (3, 300, 256, 5, 64, torch.bfloat16),    # SegFormer stage 3
    (2, 256, 64, 2, 32, torch.float16),      # MiT-B0 head dim
    (2, 200, 200, 3, 64, torch.bfloat16),    # DOFA-like: keys = queries, key rows padded to 64
])
def test_grouped_attention_wgrads(cuda, b, n, nk, heads, d, dtype):
    """dV = P^T.dO and dK = dS^T.q of all heads as ONE grouped batched-wgrad launch each: same sums as the per-head launches
    (fp32 atomics: compared with a tolerance) and as torch"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(n * 3 + nk)
    c = heads * d
    lp = (nk + 63) // 64 * 64 if nk == n else (nk + 15) // 16 * 16
    do = (torch.randn(b, 1, n, c, generator=g) * 0.3).to(dtype).cuda()
    p = torch.zeros(b, 1, n, heads * lp, dtype=dtype).cuda()
    p.view(b, n, heads, lp)[..., :nk] = torch.rand(b, n, heads, nk, generator=g).to(dtype).cuda()
    dkv_loop = torch.zeros(b, lp, 2 * c, device="cuda")
    dkv_grp = torch.zeros_like(dkv_loop)
    for hd in range(heads):
        ops.conv2d_wgrad([do[..., hd * d:(hd + 1) * d]], p[..., hd * lp:(hd + 1) * lp], 1, 1, 0, 0, dkv_loop[:, :, c + hd * d:c + (hd + 1) * d])
    ops.conv2d_wgrad([do[..., 0:d]], p[..., 0:lp], 1, 1, 0, 0, dkv_grp[:, :, c:c + d], groups=(heads, d, lp, d))
    torch.cuda.synchronize()
    ref = torch.einsum("bnhk,bnhd->bkhd", p.view(b, n, heads, lp).float(), do.view(b, n, heads, d).float()).reshape(b, lp, c)
    assert ((dkv_grp[:, :, c:] - ref).abs().max() / ref.abs().max()).item() < 2e-3
    assert ((dkv_grp - dkv_loop).abs().max() / ref.abs().max()).item() < 1e-5
    assert not dkv_grp[:, :, :c].any()  # the K half was not touched
    with pytest.raises(ValueError):  # grouped products are batched 1x1 only
        ops.conv2d_wgrad([do[..., 0:d]], p[..., 0:lp], 1, 1, 0, 0, dkv_grp[0, :, c:c + d], groups=(heads, d, lp, d))
