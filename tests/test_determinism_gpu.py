"""GPU: every cross-block reduction of the library is bit-reproducible.

With the per-device workspace registered (gdl_set_workspace; the Python binding does it on first use) the BatchNorm sums,
parameter-gradient sums, loss statistics, gradient norm and the pixel-split weight gradients are formed in a fixed order:
two identical launches must agree bit for bit, and the tickets / turnstiles must be back at zero afterwards.  The values
themselves are checked against float64 sums of the same 16-bit operands.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _twice(fn):
    a = fn()
    b = fn()
    return a, b


def _tickets_at_rest():
    from gdl_b200 import _lib
    ws = _lib.workspace_tensor()
    torch.cuda.synchronize()
    return int(ws[:256 * 1024].view(torch.int32).abs().sum().item()) == 0


def test_workspace_is_registered(cuda):
    from gdl_b200 import ops
    assert ops.deterministic()
    assert _tickets_at_rest()


@pytest.mark.parametrize("m,c", [(4 * 64 * 64, 64), (2 * 33 * 17, 24), (8 * 128 * 128, 16), (4 * 16 * 16, 2048), (300000, 256)])
def test_bn_stats_reproducible(cuda, m, c):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(m + c)
    x = (torch.randn(m, c, generator=g) * 3 + 1).to(torch.bfloat16).cuda().view(1, 1, m, c)
    piv = torch.randn(c, generator=g).cuda()

    def run():
        s = torch.empty(2 * c, device="cuda")
        ops.bn_stats(x, s, piv)
        return s
    a, b = _twice(run)
    assert torch.equal(a, b)
    d = x.view(m, c).double() - piv.double()
    ref = torch.cat([d.sum(0), (d * d).sum(0)])
    assert torch.allclose(a.double(), ref, rtol=2e-5, atol=2e-3 * (m ** 0.5))
    assert _tickets_at_rest()


@pytest.mark.parametrize("n,h,w,c,nsrc,up", [(4, 32, 32, 64, 3, True), (2, 17, 9, 16, 1, False), (8, 64, 64, 128, 2, False),
                                             (32, 16, 16, 1024, 6, False), (2, 12, 20, 32, 2, "first")])
def test_grad_gather_sums_reproducible(cuda, n, h, w, c, nsrc, up):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(n * h + c)
    srcs = [(torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda(), 0) for _ in range(nsrc)]
    if up:
        srcs.append((torch.randn(n, 2 * h, 2 * w, c, generator=g).to(torch.bfloat16).cuda(), 1))
    if up == "first":  # a 2x2-pooled source that is not the last one: the generic kernel variant
        srcs.insert(0, (torch.randn(n, 2 * h, 2 * w, c, generator=g).to(torch.bfloat16).cuda(), 1))
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    y = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    mean, invstd = torch.randn(c, generator=g).cuda() * 0.1, (torch.rand(c, generator=g) + 0.5).cuda()

    def run():
        gg = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
        s = torch.empty(2 * c, device="cuda")
        ops.grad_gather(srcs, (n, h, w, c), torch.bfloat16, y=y, x=x, mean=mean, invstd=invstd, g=gg, sums=s)
        return gg, s
    (g1, s1), (g2, s2) = _twice(run)
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    want = torch.zeros(n, h, w, c, dtype=torch.float32, device="cuda")
    for t_, mode in srcs:
        want += t_.float() if mode == 0 else t_.float().view(n, h, 2, w, 2, c).sum(dim=(2, 4))
    want = torch.where(y.float() > 0, want, torch.zeros_like(want))
    assert (g1.float() - want).abs().max() <= 0.02 * want.abs().max() + 1e-3  # one 16-bit rounding of the sum
    gd = g1.double().view(-1, c)
    xh = (x.double().view(-1, c) - mean.double()) * invstd.double()
    ref = torch.cat([gd.sum(0), (gd * xh).sum(0)])
    assert torch.allclose(s1.double(), ref, rtol=1e-4, atol=1e-2 * (n * h * w) ** 0.5)
    assert _tickets_at_rest()


@pytest.mark.parametrize("k,binary", [(5, False), (1, True), (19, False)])
def test_loss_statistics_reproducible(cuda, k, binary):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(k)
    logits = torch.randn(2, 64, 48, k, generator=g).cuda()
    t = torch.randint(0, max(k, 2), (2, 64, 48), generator=g).cuda()
    spec = ops.LossSpec(1.0, 0.5, label_smoothing=0.1, ignore_index=-100)
    (c1, s1), (c2, s2) = _twice(lambda: ops.seg_loss_fwd(logits, t, spec))
    assert torch.equal(c1, c2) and torch.equal(s1, s2)
    assert _tickets_at_rest()


def test_grad_norm_reproducible(cuda):
    from gdl_b200 import ops
    g = torch.randn(3_000_017, generator=torch.Generator().manual_seed(3)).cuda()

    def run():
        scratch, scale = torch.empty(1, device="cuda"), torch.empty(1, device="cuda")
        ops.grad_clip_coef(g, 1.0, scratch, scale)
        return scratch, scale
    (a, sa), (b, sb) = _twice(run)
    assert torch.equal(a, b) and torch.equal(sa, sb)
    assert abs(a.item() - g.double().pow(2).sum().item()) < 1e-4 * a.item()
    assert _tickets_at_rest()


@pytest.mark.parametrize("m,c", [(4 * 1024, 64), (4 * 256, 512), (1500, 320)])
def test_layernorm_param_grads_reproducible(cuda, m, c):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(m)
    x = torch.randn(m, c, generator=g).cuda()
    gy = torch.randn(m, c, generator=g).to(torch.bfloat16).cuda()
    gamma, beta = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    _, stats = ops.layernorm_fwd(x, gamma, beta, 1e-6, torch.bfloat16)

    def run():
        pg = torch.zeros(2, c, device="cuda")
        dx32, _ = ops.layernorm_bwd(gy, x, stats, gamma, pgrads=pg)
        return dx32, pg
    (d1, p1), (d2, p2) = _twice(run)
    assert torch.equal(d1, d2) and torch.equal(p1, p2)
    xh = (x.double() - stats[0].double().unsqueeze(1)) * stats[1].double().unsqueeze(1)
    ref = torch.stack([(gy.double() * xh).sum(0), gy.double().sum(0)])
    assert torch.allclose(p1.double(), ref, rtol=1e-4, atol=1e-2 * m ** 0.5)
    assert _tickets_at_rest()


def test_dwconv_and_layerscale_param_grads_reproducible(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(9)
    n, h, w, c = 4, 32, 32, 256
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    wt, bias = torch.randn(c, 3, 3, generator=g).cuda() * 0.2, torch.randn(c, generator=g).cuda() * 0.1
    y, pre = ops.dwconv3x3_gelu_fwd(x, wt, bias)
    dy = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()

    def run():
        pg = torch.zeros(c, 10, device="cuda")
        dx = ops.dwconv3x3_gelu_bwd(dy, pre, x, wt, pg)
        return dx, pg
    (d1, p1), (d2, p2) = _twice(run)
    assert torch.equal(d1, d2) and torch.equal(p1, p2)
    m, cc = 4 * 1297, 768
    gs = torch.randn(m, cc, generator=g).cuda()
    u = torch.randn(m, cc, generator=g).to(torch.bfloat16).cuda()
    gamma = torch.full((cc,), 1e-2).cuda()

    def run2():
        dg = torch.zeros(cc, device="cuda")
        du = ops.layerscale_bwd(gs, u, gamma, dg)
        return du, dg
    (u1, g1), (u2, g2) = _twice(run2)
    assert torch.equal(u1, u2) and torch.equal(g1, g2)
    assert torch.allclose(g1.double(), (gs.double() * u.double()).sum(0), rtol=1e-4, atol=1e-1)
    assert _tickets_at_rest()


# (N, H, W, [source channels], Cout, R): generic kernel with a pixel split, the halo variant, the row-streaming kernel for
# narrow outputs (several slabs / one slab / ragged width), a pointwise GEMM, a virtual concat
WGRAD_CASES = [
    (8, 64, 64, [128], 256, 3),
    (4, 128, 128, [64], 128, 3),
    (4, 128, 128, [64, 128], 64, 3),
    (2, 256, 256, [64], 32, 3),
    (3, 70, 200, [192], 16, 3),
    (16, 64, 64, [256], 64, 1),
    (4, 32, 32, [512, 256, 64], 256, 3),
]


# small shapes that still split the pixel range (also executed on the CPU functional model, tests/hostemu)
WGRAD_CASES_SMALL = [
    (2, 32, 64, [64], 128, 3),      # halo variant, 3 accumulators per unit
    (2, 16, 128, [64, 64], 32, 3),  # row-streaming kernel, two slabs
    (2, 48, 48, [32], 48, 1),       # pointwise
]


@pytest.mark.parametrize("n,h,w,cins,cout,r", WGRAD_CASES_SMALL)
def test_wgrad_reproducible_small(cuda, n, h, w, cins, cout, r):
    test_wgrad_reproducible(cuda, n, h, w, cins, cout, r)


@pytest.mark.parametrize("n,h,w,cins,cout,r", WGRAD_CASES)
def test_wgrad_reproducible(cuda, n, h, w, cins, cout, r):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(h * w + cout)
    srcs = [torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda() for c in cins]
    dy = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16).cuda()
    ctot, pad = sum(cins), r // 2

    def run():
        dw = torch.zeros(cout, r * r * ctot, device="cuda")
        ops.conv2d_wgrad(srcs, dy, r, r, pad, pad, dw)
        return dw
    a = run()
    for _ in range(3):
        assert torch.equal(a, run())
    x = torch.cat([s.float() for s in srcs], dim=3).permute(0, 3, 1, 2)
    wref = torch.nn.grad.conv2d_weight(x, (cout, ctot, r, r), dy.float().permute(0, 3, 1, 2), padding=pad)
    ref = wref.permute(0, 2, 3, 1).reshape(cout, -1)
    assert (a - ref).norm() / ref.norm() < 2e-3
    assert _tickets_at_rest()


def test_batched_wgrad_reproducible(cuda):
    """attention dV = P^T.dO per image: each image's pixel split is ordered on its own turnstile"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(11)
    b, nq, nk, d = 4, 4096, 256, 64
    p = torch.rand(b, 1, nq, nk, generator=g).to(torch.bfloat16).cuda()
    do = torch.randn(b, 1, nq, d, generator=g).to(torch.bfloat16).cuda()

    def run():
        dv = torch.zeros(b, nk, d, device="cuda")
        ops.conv2d_wgrad([do], p, 1, 1, 0, 0, dv)
        return dv
    a = run()
    for _ in range(3):
        assert torch.equal(a, run())
    ref = torch.einsum("bqk,bqd->bkd", p[:, 0].float(), do[:, 0].float())
    assert (a - ref).norm() / ref.norm() < 2e-3
    assert _tickets_at_rest()


def test_atomic_paths_still_agree(cuda):
    """option deterministic = 0 selects the fp32-atomic paths: same values up to summation order"""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 1, 50000, 64, generator=g).to(torch.bfloat16).cuda()
    s1 = torch.empty(128, device="cuda")
    ops.bn_stats(x, s1)
    ops.set_option("deterministic", 0)
    try:
        s0 = torch.empty(128, device="cuda")
        ops.bn_stats(x, s0)
    finally:
        ops.set_option("deterministic", 1)
    assert torch.allclose(s0, s1, rtol=1e-4, atol=0.5)
