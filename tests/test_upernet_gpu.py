"""GPU: MultiLevelNeck + UperNet + FCN/1x1 heads (the trainable half of the DOFA configuration) against the
reference-pinned oracle, same bar as the other model tests (<= 2.5x / 3x the autocast reference's deviation)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


@pytest.mark.parametrize("e,ch,hw,img,dtype", [(96, 64, 12, 168, torch.bfloat16), (768, 256, 18, 256, torch.bfloat16),
                                               (128, 64, 12, 168, torch.float16)])
def test_upernet_train_step_parity(cuda, e, ch, hw, img, dtype):
    from gdl_b200.models.upernet import UperNetSegmentor
    from oracle import upernet as ou
    k = 5
    torch.manual_seed(0)
    prod = UperNetSegmentor(e, ch, k, compute_dtype=dtype).cuda().train()
    with torch.no_grad():
        for _, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))

    def sd_copy():
        return {n: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n else v.clone())
                for n, v in prod.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(4, e, hw, hw, generator=g).cuda() for _ in range(4)]
    t = torch.randint(0, k, (4, img, img), generator=g).cuda()

    # fp16 needs loss scaling (Lightning's "16-mixed" GradScaler): 1/(B*H*W)-sized logit gradients are subnormal in
    # half precision.  The scale is a power of two, so un-scaling the compared gradients is exact.
    scale = 1024.0 if dtype == torch.float16 else 1.0

    def loss_of(o, a):
        return scale * (F.cross_entropy(o.float(), t) + 0.4 * F.cross_entropy(a.float(), t))
    sd = sd_copy()
    ro, ra = ou.upernet_forward(sd, feats, (img, img), training=True)
    loss_of(ro, ra).backward()
    sd_ac = sd_copy()
    with torch.autocast("cuda", dtype=dtype):
        ao, aa = ou.upernet_forward(sd_ac, feats, (img, img), training=True)
    loss_of(ao, aa).backward()
    o, a = prod(feats, (img, img))
    assert o.shape == ro.shape and a.shape == ra.shape
    loss_of(o, a).backward()
    for name, got, ref, ac in (("out", o, ro, ao), ("aux", a, ra, aa)):
        ep, ea = _rel(got, ref), _rel(ac, ref)
        print(f"[{e} {dtype}] {name} logits rel err: product {ep:.4f}, autocast reference {ea:.4f}")
        assert ep < max(2.5 * ea, 5e-3)
    rows = [(n, _rel(p.grad, sd[n].grad), _rel(sd_ac[n].grad, sd[n].grad)) for n, p in prod.named_parameters()
            if sd[n].grad.abs().max() > 1e-9]
    print(f"[{e} {dtype}] worst grad err ratio vs autocast: {max(r[1] / max(r[2], 2e-3) for r in rows):.2f}")
    for n, ep, ea in rows:
        # the PPM branches normalise 4*s*s samples per channel (16 for the 2x2 bin): one ReLU unit whose sign differs
        # between two 16-bit roundings moves that layer's gradient by ~10 % (the autocast reference shows the same
        # scatter from seed to seed), so those layers get a wider per-parameter bound and are covered by the global one
        loose = "psp_modules" in n
        assert ep < max((6.0 if loose else 3.0) * ea, 0.15 if loose else 2e-2), f"{n}: product {ep:.4f} vs autocast {ea:.4f}"
    flat = lambda d: torch.cat([d[n].flatten().float() for n, _, _ in rows])  # noqa: E731
    gp = flat({n: p.grad for n, p in prod.named_parameters()})
    gr, ga = flat({n: sd[n].grad for n in sd if sd[n].grad is not None}), flat({n: sd_ac[n].grad for n in sd_ac if sd_ac[n].grad is not None})
    ep, ea = _rel(gp, gr), _rel(ga, gr)
    print(f"[{e} {dtype}] all gradients: product {ep:.4f}, autocast reference {ea:.4f}")
    assert ep < max(2.0 * ea, 1e-2)


def test_adaptive_pool_and_add_kernels(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 18, 18, 64, generator=g).bfloat16().cuda()
    for s in (1, 2, 3, 6):
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        ref = F.adaptive_avg_pool2d(xr, s)
        y = ops.adaptive_avgpool_fwd(x, s)
        assert _rel(y, ref.permute(0, 2, 3, 1)) < 2 ** -8
        dy = torch.randn(2, s, s, 64, generator=g).bfloat16().cuda()
        ref.backward(dy.float().permute(0, 3, 1, 2))
        assert _rel(ops.adaptive_avgpool_bwd(dy, 18, 18), xr.grad.permute(0, 2, 3, 1)) < 2 ** -7
    # more bins than pixels (PPM scale 6 on the 3x3 map of a 112-pixel DOFA tile), non-square ratios
    for hw, s in ((3, 6), (5, 6), (7, 3)):
        xs = torch.randn(2, hw, hw, 64, generator=g).bfloat16().cuda()
        xr = xs.float().permute(0, 3, 1, 2).requires_grad_(True)
        ref = F.adaptive_avg_pool2d(xr, s)
        assert _rel(ops.adaptive_avgpool_fwd(xs, s), ref.permute(0, 2, 3, 1)) < 2 ** -8
        dy = torch.randn(2, s, s, 64, generator=g).bfloat16().cuda()
        ref.backward(dy.float().permute(0, 3, 1, 2))
        assert _rel(ops.adaptive_avgpool_bwd(dy, hw, hw), xr.grad.permute(0, 2, 3, 1)) < 2 ** -7
    b = torch.randn(2, 18, 18, 64, generator=g).bfloat16().cuda()
    assert torch.equal(ops.add_nhwc(x, b), (x.float() + b.float()).bfloat16())
