"""GPU parity of the MixTransformer/SegFormer HBM-bound kernels against torch fp32 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("c", [32, 64, 160, 320, 512, 768])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(cuda, c, xdt):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(c)
    m = 300
    x = (torch.randn(3, m // 3, c, generator=g) * 2 + 0.5).to(xdt).cuda()
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).cuda()
    beta = (0.1 * torch.randn(c, generator=g)).cuda()
    xr = x.float().detach().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (c,), gr, br, 1e-6)
    y, stats = ops.layernorm_fwd(x, gamma, beta, 1e-6, torch.float32)
    assert _relerr(y, ref) < 1e-5
    y16, _ = ops.layernorm_fwd(x, gamma, beta, 1e-6, BF)
    assert _relerr(y16, ref) < 2 ** -8
    dy = (torch.randn(3, m // 3, c, generator=g)).to(BF).cuda()
    add = torch.randn(3, m // 3, c, generator=g).cuda()
    ref.backward(dy.float())
    pg = torch.zeros(2, c, device="cuda")
    dx32, dx16 = ops.layernorm_bwd(dy, x, stats, gamma, add=add, want32=True, dtype16=BF, pgrads=pg)
    assert _relerr(dx32, xr.grad + add) < 1e-4
    assert _relerr(dx16, xr.grad + add) < 2 ** -8
    assert _relerr(pg[0], gr.grad) < 1e-4 and _relerr(pg[1], br.grad) < 1e-4


@pytest.mark.parametrize("length,lpad", [(256, 256), (4, 16), (9, 16), (64, 64), (1000, 1008)])
def test_softmax_fwd_bwd(cuda, length, lpad):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(length)
    s = (torch.randn(2, 3, 50, lpad, generator=g) * 3).to(BF).cuda()
    scale = 0.125
    sr = s[..., :length].float().detach().requires_grad_(True)
    ref = torch.softmax(sr * scale, -1)
    p = ops.softmax_fwd(s, scale, length)
    assert _relerr(p[..., :length], ref) < 2 ** -8
    assert p[..., length:].abs().max() == 0 if lpad > length else True
    dp = torch.randn(2, 3, 50, lpad, generator=g).to(BF).cuda()
    # reference backward through the bf16-rounded p the kernel saved
    pk = p[..., :length].float()
    dpf = dp[..., :length].float()
    want = scale * pk * (dpf - (dpf * pk).sum(-1, keepdim=True))
    ds = ops.softmax_bwd(p, dp, scale, length)
    assert _relerr(ds[..., :length], want) < 2 ** -7
    assert ds[..., length:].abs().max() == 0 if lpad > length else True


@pytest.mark.parametrize("c,h,w", [(256, 16, 16), (512, 8, 12), (2048, 4, 4), (128, 32, 32)])
def test_dwconv_gelu_fwd_bwd(cuda, c, h, w):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(c + h)
    n = 2
    x = torch.randn(n, h, w, c, generator=g).to(BF).cuda()
    wt = (torch.randn(c, 1, 3, 3, generator=g) * 0.3).cuda()
    b = (torch.randn(c, generator=g) * 0.1).cuda()
    xr = x.float().permute(0, 3, 1, 2).detach().requires_grad_(True)
    wr, brr = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre_ref = F.conv2d(xr, wr, brr, padding=1, groups=c)
    ref = F.gelu(pre_ref)
    y, pre = ops.dwconv3x3_gelu_fwd(x, wt.view(c, 9).contiguous(), b)
    assert _relerr(pre, pre_ref.permute(0, 2, 3, 1)) < 2 ** -8
    assert _relerr(y, ref.permute(0, 2, 3, 1)) < 2 ** -7
    dy = torch.randn(n, h, w, c, generator=g).to(BF).cuda()
    ref.backward(dy.float().permute(0, 3, 1, 2))
    pg = torch.zeros(c, 10, device="cuda")
    dx = ops.dwconv3x3_gelu_bwd(dy, pre, x, wt.view(c, 9).contiguous(), pg)
    assert _relerr(dx, xr.grad.permute(0, 2, 3, 1)) < 0.02
    assert _relerr(pg[:, :9], wr.grad.view(c, 9)) < 0.02
    assert _relerr(pg[:, 9], brr.grad) < 0.02


@pytest.mark.parametrize("hi,wi,ho,wo,c,dt", [
    (16, 16, 128, 128, 64, torch.bfloat16), (32, 32, 128, 128, 24, torch.bfloat16), (64, 64, 128, 128, 8, torch.bfloat16),
    (128, 128, 512, 512, 5, torch.float32), (36, 36, 18, 18, 16, torch.bfloat16), (18, 18, 512, 512, 5, torch.float32),
    (9, 13, 40, 31, 8, torch.float32),
])
def test_bilinear_fwd_bwd(cuda, hi, wi, ho, wo, c, dt):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(hi * 7 + wo)
    n = 2
    x = torch.randn(n, hi, wi, c, generator=g).to(dt).cuda()
    xr = x.float().permute(0, 3, 1, 2).detach().requires_grad_(True)
    ref = F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=False)
    y = ops.bilinear_fwd(x, ho, wo)
    tol = 1e-5 if dt == torch.float32 else 2 ** -8
    assert _relerr(y, ref.permute(0, 2, 3, 1)) < tol
    dy = torch.randn(n, ho, wo, c, generator=g).to(dt).cuda()
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dx = ops.bilinear_bwd(dy, hi, wi)
    assert _relerr(dx, xr.grad.permute(0, 2, 3, 1)) < (1e-5 if dt == torch.float32 else 2 ** -7)


@pytest.mark.parametrize("n,ho,wo,c,lows,dt", [
    (2, 32, 32, 64, [(4, 4), (8, 8), (16, 16)], torch.bfloat16),     # the SegFormer decoder: strides 32 / 16 / 8 onto 4
    (1, 24, 40, 16, [(3, 5), (12, 20)], torch.bfloat16),             # two sources, non-square
    (3, 16, 16, 8, [(5, 7)], torch.float16),                         # one source, non-integer scale
    (2, 8, 8, 32, [], torch.bfloat16),                               # no resized source: a copy of the base
])
def test_bilinear_sum_fwd(cuda, n, ho, wo, c, lows, dt):
    """out = base + sum_i resize(src_i) (gdl_bilinear_sum_fwd): fp32 sum in the given order, ONE rounding — against
    F.interpolate in fp32 on the same 16-bit operands; channel slices of wider tensors (row stride > C) as inputs and output;
    and the identity that makes the folded SegFormer decoder valid: a channel mix applied before the resizes equals the
    mix applied to the concatenated resized maps."""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(ho * 3 + c)
    wide = torch.randn(n, ho, wo, c + 8, generator=g).to(dt).cuda()
    base = wide[..., 8:]                                             # row stride c + 8
    srcs = [torch.randn(n, h, w, c, generator=g).to(dt).cuda() for h, w in lows]
    want = base.float()
    for t in srcs:
        want = want + F.interpolate(t.float().permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear",
                                    align_corners=False).permute(0, 2, 3, 1)
    y = ops.bilinear_sum_fwd(base, srcs)
    ulp = 2 ** -8 if dt == torch.bfloat16 else 2 ** -11
    assert (y.float() - want).abs().max() <= ulp * want.abs().max() + 1e-6      # one rounding of the fp32 sum
    out_wide = torch.zeros(n, ho, wo, c + 16, dtype=dt, device="cuda")
    ops.bilinear_sum_fwd(base, srcs, out=out_wide[..., 16:])
    assert torch.equal(out_wide[..., 16:], y) and not out_wide[..., :16].any()
    if not srcs:
        assert torch.equal(y, base)
        return
    # linearity: mix(concat(resize(s_i))) == sum_i resize(mix_i(s_i))   (mix = a 1x1 conv without bias)
    k = len(srcs)
    w = torch.randn(c, k * c, generator=g).cuda() / (k * c) ** 0.5
    ups = [F.interpolate(t.float().permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear", align_corners=False) for t in srcs]
    ref = F.conv2d(torch.cat(ups, 1), w.view(c, k * c, 1, 1)).permute(0, 2, 3, 1)
    mixed = [F.conv2d(t.float().permute(0, 3, 1, 2), w[:, i * c:(i + 1) * c].reshape(c, c, 1, 1)).permute(0, 2, 3, 1)
             .contiguous().to(dt) for i, t in enumerate(srcs)]
    got = ops.bilinear_sum_fwd(torch.zeros(n, ho, wo, c, dtype=dt, device="cuda"), mixed)
    assert _relerr(got, ref) < 4 * ulp
    with pytest.raises(ValueError):
        ops.bilinear_sum_fwd(base, [srcs[0][..., :c - 8].contiguous()] if c > 8 else [srcs[0].float()])


def test_cast_f32(cuda):
    from gdl_b200 import ops
    x = torch.randn(1000, 7, device="cuda")
    assert torch.equal(ops.cast_f32(x, BF), x.to(BF))
