"""GPU: DOFA-v2 encoder forward (frozen) and the DOFA segmentation model's train step against the reference-pinned
oracle (oracle/dofa.py + oracle/upernet.py) and the committed reference golden vectors.  Bars: kernels vs fp32 on the
same 16-bit operands; whole encoder <= 2.5x the deviation of the oracle under torch.autocast."""
import pathlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).parent / "golden"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def _load(prod, sd):
    prod.load_state_dict(sd)
    for p in prod.parameters():
        p.requires_grad_(False)
    return prod.cuda()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_dofa_encoder_matches_reference_golden(cuda, dtype):
    """The golden file holds the REFERENCE DOFAv2's own outputs (oracle/make_golden.py::dofa_golden)."""
    from gdl_b200.models.dofa import DOFAv2
    from oracle import dofa as od
    g = torch.load(GOLD / "dofa_golden.pt")
    sd = od.init_state_dict(192, 4, 112, seed=2, ls_init=0.5)
    prod = _load(DOFAv2("dofa_base", 112, 14, 192, 4, 3, out_indices=[1, 2, 3], compute_dtype=dtype), sd)
    x, wl = g["x"].cuda(), g["wavelengths"].cuda()
    got = prod(x, wl)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.autocast("cuda", dtype=dtype):
        ac = od.dofa_forward(sdc, x, wl, 192, 4, 3, out_indices=(1, 2, 3))
    for i, (a, c, want) in enumerate(zip(got, ac, g["feats"])):
        want = want.cuda()
        assert a.shape == want.shape
        ep, ea = _rel(a, want), _rel(c, want)
        print(f"[{dtype}] dofa feature {i}: product {ep:.5f}, autocast oracle {ea:.5f}")
        assert ep < max(2.5 * ea, 4e-3)


def test_dofa_base_full_width_against_oracle(cuda):
    """dofa_base width (768, 12 heads) at 224x224 (257 tokens) and 6 bands, 12 blocks, LayerScale 0.3."""
    from gdl_b200.models.dofa import create_dofa_base
    from oracle import dofa as od
    sd = od.init_state_dict(768, 12, 224, seed=5, ls_init=0.3)
    prod = _load(create_dofa_base(224), sd)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 224, 224, generator=gen).cuda()
    wl = torch.tensor([0.49, 0.56, 0.665, 0.842, 1.61, 2.19]).cuda()
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        want = od.dofa_forward(sdc, x, wl)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = od.dofa_forward(sdc, x, wl)
    got = prod(x, wl)
    assert len(got) == 4
    for i, (a, c, w) in enumerate(zip(got, ac, want)):
        ep, ea = _rel(a, w), _rel(c, w)
        print(f"dofa_base tap {i}: product {ep:.5f}, autocast oracle {ea:.5f}")
        assert a.shape == w.shape == (2, 768, 16, 16) and ep < max(2.5 * ea, 4e-3)


def test_dofa_full_token_count_kernels(cuda):
    """The 512x512 configuration has 1297 tokens per image: odd GEMM width, padded softmax rows (Lp = 1312)."""
    from gdl_b200.models.dofa import _mha
    torch.manual_seed(0)
    b, n, heads, d = 2, 1297, 12, 64
    c = heads * d
    qkv = (0.5 * torch.randn(b * n, 3 * c)).bfloat16().cuda()
    o = _mha(qkv, b, n, heads, c, torch.bfloat16)
    q, k, v = (t.view(b, n, heads, d).permute(0, 2, 1, 3) for t in qkv.float().split(c, dim=1))
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(b * n, c)
    err = _rel(o, ref)
    print(f"attention, 1297 tokens: rel err {err:.5f}")
    assert err < 1e-2  # P is rounded to bf16 before P.V (as under autocast)


def test_gelu_and_layerscale_epilogues(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 1, 300, 128, generator=g).bfloat16().cuda()
    w = (0.1 * torch.randn(200, 128, generator=g)).bfloat16().cuda()
    bias = torch.randn(200, generator=g).cuda()
    gamma = torch.randn(200, generator=g).cuda()
    res = torch.randn(1, 1, 300, 200, generator=g).cuda()
    acc = x.float().view(300, 128) @ w.float().t() + bias
    try:
        for mode in (0, 1, 2):  # every epilogue variant of the GEMM kernel
            ops.set_option("conv_epilogue", mode)
            y = ops.conv2d_fwd([x], w, 200, 1, 1, 0, 0, bias=bias, gelu=True, out_dtype=torch.float32)
            assert _rel(y.view(300, 200), F.gelu(acc)) < 1e-5, mode
            y = ops.conv2d_fwd([x], w, 200, 1, 1, 0, 0, bias=bias, oscale=gamma, residual=res, out_dtype=torch.float32)
            assert _rel(y.view(300, 200), acc * gamma + res.view(300, 200)) < 1e-5, mode
            y = ops.conv2d_fwd([x], w, 200, 1, 1, 0, 0, bias=bias, gelu=True)
            assert _rel(y.view(300, 200), F.gelu(acc)) < 2 ** -8, mode
    finally:
        ops.set_option("conv_epilogue", 2)


def test_vit_token_kernels(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(1)
    patch = torch.randn(3, 49, 96, generator=g).bfloat16().cuda()
    pos = torch.randn(50, 96, generator=g).cuda()
    cls = torch.randn(96, generator=g).cuda()
    tok = ops.vit_assemble_tokens(patch, pos, cls)
    assert tok.dtype == torch.float32 and torch.equal(tok[:, 0], cls.expand(3, 96))
    assert torch.equal(tok[:, 1:], patch.float() + pos[1:])
    feat = ops.vit_extract_feature(tok, torch.bfloat16)
    assert torch.equal(feat, tok[:, 1:].bfloat16())


def test_dofa_segmentation_train_step(cuda, img=224, bands=3, batch=4, wavelengths=(0.665, 0.56, 0.49)):
    """DOFASegmentationModel (frozen encoder) forward + backward through autograd vs the oracle on the same maps."""
    from gdl_b200.models.dofa import DOFASegmentationModel
    from oracle import dofa as od, upernet as ou
    torch.manual_seed(0)
    k = 5
    m = DOFASegmentationModel("dofa_base", (img, img), ["encoder"], k).cuda().train()
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "ls1" in n_ or "ls2" in n_:
                p.fill_(0.3)
            elif p.dim() == 1 and not n_.startswith("encoder."):
                p.add_(0.1 * torch.randn_like(p))
    assert not any(p.requires_grad for p in m.encoder.parameters())
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(batch, bands, img, img, generator=gen).cuda()
    wl = torch.tensor(list(wavelengths)).cuda()
    t = torch.randint(0, k, (batch, img, img), generator=gen).cuda()

    def sd_copy():
        return {n: (v.detach().clone().requires_grad_(True)
                    if v.is_floating_point() and "running" not in n and not n.startswith("encoder.") else v.clone())
                for n, v in m.state_dict().items()}

    def oracle(sd):
        enc = {n[len("encoder."):]: v for n, v in sd.items() if n.startswith("encoder.")}
        with torch.no_grad():
            feats = od.dofa_forward(enc, x, wl)
        return ou.upernet_forward({n: v for n, v in sd.items() if not n.startswith("encoder.")}, feats, (img, img), training=True)

    def loss_of(o, a):
        return F.cross_entropy(o.float(), t) + 0.4 * F.cross_entropy(a.float(), t)
    sd = sd_copy()
    ro, ra = oracle(sd)
    loss_of(ro, ra).backward()
    sd_ac = sd_copy()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ao, aa = oracle(sd_ac)
    loss_of(ao, aa).backward()
    out = m(x, wl)
    loss_of(out.out, out.aux).backward()
    for name, got, ref, ac in (("out", out.out, ro, ao), ("aux", out.aux, ra, aa)):
        ep, ea = _rel(got, ref), _rel(ac, ref)
        print(f"dofa seg {name} logits rel err: product {ep:.4f}, autocast oracle {ea:.4f}")
        assert ep < max(2.5 * ea, 5e-3)
    rows = [(n, _rel(p.grad, sd[n].grad), _rel(sd_ac[n].grad, sd[n].grad)) for n, p in m.named_parameters()
            if p.requires_grad and sd[n].grad.abs().max() > 1e-9]
    assert rows and all(p.grad is None for n, p in m.named_parameters() if n.startswith("encoder."))
    print(f"dofa seg worst grad err ratio vs autocast: {max(r[1] / max(r[2], 2e-3) for r in rows):.2f}")
    for n, ep, ea in rows:
        assert ep < max(3.0 * ea, 2e-2), f"{n}: product {ep:.4f} vs autocast {ea:.4f}"
    # eval path returns the same structure
    m.eval()
    with torch.no_grad():
        ev = m(x, wl)
    assert ev.out.shape == (batch, k, img, img) and ev.aux.shape == (batch, k, img, img)
