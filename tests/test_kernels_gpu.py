"""GPU parity tests of every C-ABI kernel against torch fp32 / the oracle on the same seeded inputs.

Tolerances: tensor-core kernels accumulate in fp32 from 16-bit operands, so against an fp32
reference evaluated ON THE SAME 16-bit-rounded operands the error is accumulation-order noise
(<= 2e-3 of the output scale is asserted, ~1e-6 observed).  Outputs stored as bf16 carry one
rounding (2^-9 relative).  Integer outputs (argmax, pooling indices) are bit-exact.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _rand(shape, g, scale=0.5, dtype=BF):
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


# ----------------------------------------------------------------------------------------------
# tensor-core convolution
# ----------------------------------------------------------------------------------------------
CONV_CASES = [
    # n, h, w, [src channels], cout, r, pad
    (1, 1, 128, [64], 16, 1, 0),
    (1, 1, 1000, [256], 256, 1, 0),
    (2, 16, 16, [320], 320, 1, 0),
    (1, 16, 16, [16], 5, 1, 0),
    (2, 32, 32, [64], 64, 3, 1),
    (2, 64, 64, [128], 256, 3, 1),
    (2, 32, 32, [64, 128, 64], 128, 3, 1),
    (1, 64, 64, [32], 16, 3, 1),
    (1, 64, 64, [16], 16, 3, 1),
    (1, 32, 32, [16], 5, 3, 1),
    (2, 36, 36, [64], 64, 3, 1),
    (2, 18, 18, [128], 64, 3, 1),
    (1, 32, 32, [64], 32, 7, 3),
    (1, 20, 20, [64], 32, 3, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_conv_fwd_matches_fp32_conv(cuda, case, dt):
    from gdl_b200 import ops
    n, h, w, chans, cout, r, pad = case
    g = torch.Generator().manual_seed(1234)
    srcs = [_rand((n, h, w, c), g, dtype=dt) for c in chans]
    ctot = sum(chans)
    wt = (torch.randn(cout, ctot, r, r, generator=g) / (ctot * r * r) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    wp = ops.pack_conv_weight(wt, dt)
    assert torch.equal(wp.view(cout, r, r, ctot), wt.permute(0, 2, 3, 1).contiguous().to(dt))
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2).contiguous()
    w32 = wp.view(cout, r, r, ctot).float().permute(0, 3, 1, 2).contiguous()
    ref = F.relu(F.conv2d(x, w32, bias, padding=pad)).permute(0, 2, 3, 1)
    out = ops.conv2d_fwd(srcs, wp, cout, r, r, pad, pad, out_dtype=torch.float32, bias=bias, relu=True)
    assert _relerr(out, ref) < 2e-3
    out16 = ops.conv2d_fwd(srcs, wp, cout, r, r, pad, pad)
    ref16 = F.conv2d(x, w32, None, padding=pad).permute(0, 2, 3, 1)
    tol = 2 ** -8 if dt == torch.bfloat16 else 2 ** -10
    assert _relerr(out16, ref16) < tol


def test_conv_reads_channel_slices_and_writes_strided(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(5)
    big = _rand((1, 16, 16, 128), g)
    src = big[..., 32:96]
    wt = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    wp = ops.pack_conv_weight(wt, BF)
    outbuf = torch.zeros(1, 16, 16, 192, dtype=BF, device="cuda")
    ops.conv2d_fwd([src], wp, 64, 3, 3, 1, 1, out=outbuf[..., 64:128])
    ref = F.conv2d(src.float().permute(0, 3, 1, 2), wp.view(64, 3, 3, 64).float().permute(0, 3, 1, 2), padding=1)
    assert _relerr(outbuf[..., 64:128], ref.permute(0, 2, 3, 1)) < 2 ** -8
    assert outbuf[..., :64].abs().max() == 0 and outbuf[..., 128:].abs().max() == 0


@pytest.mark.parametrize("case", [
    (1, 1, 256, [64], 64, 1, 0), (1, 1, 4096, [320], 256, 1, 0), (1, 1, 512, [16], 16, 1, 0),
    (2, 32, 32, [128], 256, 3, 1), (2, 32, 32, [64, 128, 64], 128, 3, 1), (1, 64, 64, [16], 16, 3, 1),
    (1, 64, 64, [32], 16, 3, 1), (2, 36, 36, [64], 64, 3, 1), (1, 32, 32, [64], 32, 7, 3),
])
def test_conv_wgrad_matches_autograd(cuda, case):
    from gdl_b200 import ops
    n, h, w, chans, cout, r, pad = case
    g = torch.Generator().manual_seed(99)
    srcs = [_rand((n, h, w, c), g) for c in chans]
    ctot = sum(chans)
    ho, wo = h + 2 * pad - r + 1, w + 2 * pad - r + 1
    dy = _rand((n, ho, wo, cout), g)
    dw = torch.zeros(cout, r * r * ctot, device="cuda")
    ops.conv2d_wgrad(srcs, dy, r, r, pad, pad, dw)
    x = torch.cat([t.double() for t in srcs], 3).permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.grad.conv2d_weight(x, (cout, ctot, r, r), dy.double().permute(0, 3, 1, 2).contiguous(), padding=pad)
    assert _relerr(dw.view(cout, r, r, ctot), ref.permute(0, 2, 3, 1)) < 2e-3
    # layout transform back to OIHW
    oihw = torch.empty(cout, ctot, r, r, device="cuda")
    ops.unpack_conv_wgrad(dw, oihw)
    assert torch.equal(oihw, dw.view(cout, r, r, ctot).permute(0, 3, 1, 2).contiguous())


@pytest.mark.parametrize("cin,cout,r,pad", [(64, 128, 3, 1), (256, 64, 1, 0), (32, 16, 3, 1), (64, 32, 7, 3)])
def test_conv_dgrad_through_transposed_weights(cuda, cin, cout, r, pad):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 32, 32
    wt = (torch.randn(cout, cin, r, r, generator=g) / (cin * r * r) ** 0.5).cuda()
    dy = _rand((n, h, w, cout), g)
    w16 = wt.to(BF).float()
    wtp = ops.pack_conv_weight(wt, BF, 1)
    dx = ops.conv2d_fwd([dy], wtp, cin, r, r, r - 1 - pad, r - 1 - pad, out_dtype=torch.float32)
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), w16, dy.float().permute(0, 3, 1, 2).contiguous(), padding=pad)
    assert _relerr(dx, ref.permute(0, 2, 3, 1)) < 2e-3


# ----------------------------------------------------------------------------------------------
# normalise / im2col / col2im
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c", [3, 4, 6])
def test_normalize_matches_reference_golden(cuda, c):
    """Golden vectors come from the reference's utils/tensors.py itself (oracle/make_golden.py)."""
    from pathlib import Path
    from gdl_b200 import ops
    g = torch.load(Path(__file__).parent / "golden" / "tensors_golden.pt")[f"c{c}"]
    raw = g["raw"].cuda()  # (2,C,16,16) uint8 NCHW
    y = ops.normalize_to_nhwc(raw, True, BF, 8, g["mean"].cuda(), g["std"].cuda(), 255.0)
    ref = g["standardized"].permute(0, 2, 3, 1).to(BF).cuda()
    assert torch.equal(y[..., :c], ref)  # bit-exact: same fp32 arithmetic, one bf16 rounding
    assert y[..., c:].abs().max() == 0
    y2 = ops.normalize_to_nhwc(raw.permute(0, 2, 3, 1).contiguous(), False, torch.float16, 8, g["mean"].cuda(),
                               g["std"].cuda(), 255.0)
    assert torch.equal(y2[..., :c], g["standardized"].permute(0, 2, 3, 1).half().cuda())
    # cast-only path used for batch["image"] (already standardised float NCHW)
    y3 = ops.normalize_to_nhwc(g["standardized"].cuda(), True, BF, 8)
    assert torch.equal(y3[..., :c], ref)


@pytest.mark.parametrize("c,r,stride,pad,ld", [(4, 7, 2, 3, 8), (3, 7, 2, 3, 8), (128, 3, 2, 1, 128), (256, 1, 2, 0, 256),
                                              (64, 3, 1, 1, 64)])
def test_im2col_col2im(cuda, c, r, stride, pad, ld):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(8)
    n, h, w = 2, 32, 32
    x = torch.zeros(n, h, w, ld, dtype=BF, device="cuda")
    x[..., :c] = _rand((n, h, w, c), g)
    k = r * r * c
    kpad = (k + 63) // 64 * 64
    col = ops.im2col(x, c, r, r, stride, pad, kpad)
    ho = (h + 2 * pad - r) // stride + 1
    unf = F.unfold(x[..., :c].float().permute(0, 3, 1, 2), r, padding=pad, stride=stride)  # (n, c*r*r, L), (c,r,s) order
    ref = unf.view(n, c, r * r, ho, ho).permute(0, 3, 4, 2, 1).reshape(n, ho, ho, k)
    assert torch.equal(col[..., :k].float(), ref)
    assert col[..., k:].abs().max() == 0 if kpad > k else True
    if c % 8 == 0:
        dcol = _rand((n, ho, ho, kpad), g)
        dx = ops.col2im(dcol, n, h, w, c, r, r, stride, pad)
        d = dcol[..., :k].float().view(n, ho, ho, r * r, c).permute(0, 4, 3, 1, 2).reshape(n, c * r * r, ho * ho)
        ref = F.fold(d, (h, w), r, padding=pad, stride=stride).permute(0, 2, 3, 1)
        assert _relerr(dx, ref) < 2 ** -7


# ----------------------------------------------------------------------------------------------
# batch norm forward / backward, gradient gather
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,n,h,w", [(64, 2, 16, 16), (320, 2, 8, 8), (2048, 2, 4, 4), (16, 1, 64, 64)])
def test_bn_train_forward(cuda, c, n, h, w):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(n, h, w, c, generator=g) * 1.5 + 0.7).to(BF).cuda()
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = (torch.randn(c, generator=g) * 0.1).cuda()
    rm = (torch.randn(c, generator=g) * 0.1).cuda()
    rv = (torch.rand(c, generator=g) + 0.5).cuda()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(x.float().permute(0, 3, 1, 2), rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    sums = torch.empty(2 * c, device="cuda")
    buf = torch.empty(4, c, device="cuda")
    ops.bn_stats(x, sums, rm)
    ops.bn_finalize(rm, sums, n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, buf[0], buf[1], buf[2], buf[3])
    assert torch.allclose(rm, rm_ref, atol=1e-5, rtol=1e-4)
    assert torch.allclose(rv, rv_ref, atol=1e-5, rtol=1e-4)
    xf = x.float()
    assert torch.allclose(buf[2], xf.mean((0, 1, 2)), atol=1e-4)
    assert torch.allclose(buf[3], (xf.var((0, 1, 2), unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    y = torch.empty_like(x)
    yu = torch.empty(n, 2 * h, 2 * w, c, dtype=BF, device="cuda")
    ops.bn_apply(x, buf[0], buf[1], relu=True, y=y, y_up=yu)
    ref_y = F.relu(ref).permute(0, 2, 3, 1)
    assert (y.float() - ref_y).abs().max() < 0.03
    assert torch.equal(yu, y.repeat_interleave(2, 1).repeat_interleave(2, 2))
    # residual variants
    res = _rand((n, h, w, c), g)
    ops.bn_apply(x, buf[0], buf[1], res=res, relu=True, y=y)
    assert (y.float() - F.relu(ref.permute(0, 2, 3, 1) + res.float())).abs().max() < 0.04
    ops.bn_apply(x, buf[0], buf[1], res=res, rscale=gamma, rshift=beta, relu=False, y=y)
    assert (y.float() - (ref.permute(0, 2, 3, 1) + res.float() * gamma + beta)).abs().max() < 0.06


def test_bn_eval_coeffs(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(2)
    c = 64
    gamma, beta = torch.rand(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    rm, rv = torch.randn(c, generator=g).cuda(), (torch.rand(c, generator=g) + 0.1).cuda()
    sc, sh = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_eval_coeffs(gamma, beta, rm, rv, 1e-5, sc, sh)
    x = torch.randn(4, c, generator=g).cuda()
    ref = F.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert torch.allclose(x * sc + sh, ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("c", [64, 320])
def test_relu_bn_backward_with_gather(cuda, c):
    """y = relu(bn(x)) consumed by two same-resolution consumers and one x2-upsampled consumer."""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(33)
    n, h, w = 2, 8, 8
    x16 = (torch.randn(n, h, w, c, generator=g) * 1.2 + 0.3).to(BF).cuda()
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = (torch.randn(c, generator=g) * 0.2).cuda()
    # strided gradient sources (channel slices of wider dgrad outputs)
    wide1 = _rand((n, h, w, c + 64), g)
    g1 = wide1[..., 32:32 + c]
    g2 = _rand((n, h, w, c), g)
    gu = _rand((n, 2 * h, 2 * w, c), g)

    x = x16.float().permute(0, 3, 1, 2).requires_grad_(True)
    gam, bet = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.relu(F.batch_norm(x, None, None, gam, bet, True, 0.1, 1e-5))
    up = F.interpolate(y, scale_factor=2.0, mode="nearest")
    (y * (g1.float() + g2.float()).permute(0, 3, 1, 2)).sum().backward(retain_graph=True)
    (up * gu.float().permute(0, 3, 1, 2)).sum().backward()

    sums = torch.empty(2 * c, device="cuda")
    buf = torch.empty(4, c, device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    ops.bn_stats(x16, sums, None)
    ops.bn_finalize(None, sums, n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, buf[0], buf[1], buf[2], buf[3])
    yk = torch.empty_like(x16)
    ops.bn_apply(x16, buf[0], buf[1], relu=True, y=yk)
    gbuf = torch.empty_like(x16)
    bsums = torch.empty(2 * c, device="cuda")
    ops.grad_gather([(g1, 0), (g2, 0), (gu, 1)], x16.shape, BF, y=yk, x=x16, mean=buf[2], invstd=buf[3], g=gbuf,
                    sums=bsums)
    dx = torch.empty_like(x16)
    pg = torch.empty(2, c, device="cuda")
    ops.bn_bwd_apply(gbuf, x16, buf[2], buf[3], gamma, bsums, dx, pg[0], pg[1], False)
    ref_dx = x.grad.permute(0, 2, 3, 1)
    assert _relerr(dx, ref_dx) < 0.02
    assert _relerr(pg[0], gam.grad) < 0.01
    assert _relerr(pg[1], bet.grad) < 0.01
    # gather-only (no mask, no sums) == plain sum
    plain = torch.empty_like(x16)
    ops.grad_gather([(g1, 0), (g2, 0)], x16.shape, BF, g=plain)
    assert torch.equal(plain, (g1.float() + g2.float()).to(BF))


# ----------------------------------------------------------------------------------------------
# max pool
# ----------------------------------------------------------------------------------------------
def test_maxpool_fwd_bwd_with_ties(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(4)
    n, h, w, c = 2, 16, 16, 64
    x16 = F.relu(torch.randn(n, h, w, c, generator=g)).to(BF).cuda()  # many exact-zero ties, like post-ReLU maps
    y, idx = ops.maxpool3x3s2_fwd(x16, True)
    x = x16.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref, ref_idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    dy = _rand(tuple(y.shape), g)
    dx = ops.maxpool3x3s2_bwd(dy, idx, h, w)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    assert _relerr(dx, x.grad.permute(0, 2, 3, 1)) < 2 ** -7


# ----------------------------------------------------------------------------------------------
# losses, argmax, optimizer
# ----------------------------------------------------------------------------------------------
def _loss_case(k, g, n=2, h=16, w=16):
    logits = (torch.randn(n, h, w, k, generator=g) * 2).cuda()
    if k == 1:
        t = torch.randint(0, 2, (n, h, w), generator=g).cuda()
    else:
        t = torch.randint(0, k, (n, h, w), generator=g).cuda()
    return logits, t


@pytest.mark.parametrize("k", [2, 5, 12])
@pytest.mark.parametrize("kind", ["ce", "ce_ls", "ce_ignore", "softce", "dice", "dice_smooth", "mix"])
def test_seg_loss_matches_oracle(cuda, k, kind):
    from gdl_b200 import ops
    from oracle import losses as ol
    g = torch.Generator().manual_seed(17 + k)
    logits, t = _loss_case(k, g)
    x = logits.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
    if kind == "ce":
        spec, ref = ops.LossSpec(1.0, 0.0, ignore_index=-100), F.cross_entropy(x, t)
    elif kind == "ce_ls":
        spec, ref = ops.LossSpec(1.0, 0.0, 0.1, ignore_index=-100), F.cross_entropy(x, t, label_smoothing=0.1)
    elif kind == "ce_ignore":
        t = t.clone()
        t[0, :4] = 255
        spec, ref = ops.LossSpec(1.0, 0.0, ignore_index=255), F.cross_entropy(x, t, ignore_index=255)
    elif kind == "softce":
        spec, ref = ops.LossSpec(1.0, 0.0, 0.1, True, -100), ol.soft_ce_loss(x, t, 0.1)
    elif kind == "dice":
        spec, ref = ops.LossSpec(0.0, 1.0), ol.dice_loss(x, t, "multiclass")
    elif kind == "dice_smooth":
        spec, ref = ops.LossSpec(0.0, 1.0, dice_smooth=1.0), ol.dice_loss(x, t, "multiclass", smooth=1.0)
    else:
        spec = ops.LossSpec(0.7, 0.3, ignore_index=-100)
        ref = 0.7 * F.cross_entropy(x, t) + 0.3 * ol.dice_loss(x, t, "multiclass")
    coeff, _ = ops.seg_loss_fwd(logits, t, spec)
    assert abs(coeff[0].item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    ref.backward()
    d = torch.empty_like(logits)
    ops.seg_loss_bwd(logits, t, spec, coeff, None, d)
    rg = x.grad.permute(0, 2, 3, 1)
    assert (d - rg).abs().max().item() < 1e-5 * rg.abs().max().item() + 1e-9
    # uint8 targets, 16-bit padded output (what the head backward consumes), external scale
    if kind in ("ce", "dice"):
        d16 = torch.zeros(*logits.shape[:3], 16, dtype=BF, device="cuda")
        ops.seg_loss_bwd(logits, t.to(torch.uint8), spec, coeff, torch.tensor([2.0], device="cuda"), d16)
        assert _relerr(d16[..., :k], 2 * rg) < 2 ** -8
        assert d16[..., k:].abs().max() == 0


def test_binary_dice_matches_oracle(cuda):
    from gdl_b200 import ops
    from oracle import losses as ol
    g = torch.Generator().manual_seed(5)
    logits, t = _loss_case(1, g)
    x = logits.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
    ref = ol.dice_loss(x, t.unsqueeze(1).float(), "binary")
    spec = ops.LossSpec(0.0, 1.0)
    coeff, _ = ops.seg_loss_fwd(logits, t, spec)
    assert abs(coeff[0].item() - ref.item()) < 1e-5
    ref.backward()
    d = torch.empty_like(logits)
    ops.seg_loss_bwd(logits, t, spec, coeff, None, d)
    assert (d - x.grad.permute(0, 2, 3, 1)).abs().max().item() < 1e-5 * x.grad.abs().max().item() + 1e-9


def test_argmax_bit_exact(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(6)
    logits, _ = _loss_case(5, g, 2, 32, 32)
    logits[0, 0, 0] = 1.0  # exact tie: first index wins, as torch
    assert torch.equal(ops.argmax_classes(logits), logits.softmax(3).argmax(3))
    l1, _ = _loss_case(1, g)
    assert torch.equal(ops.argmax_classes(l1), (l1[..., 0].sigmoid() > 0.5).long())


def test_adam_and_clip_match_torch(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(7)
    n = 100_003
    p0 = torch.randn(n, generator=g).cuda()
    p = p0.clone()
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    scratch = torch.zeros(2, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, generator=g).cuda() * 3
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        ops.grad_clip_coef(grad, 1.0, scratch[0:1], scratch[1:2])
        ops.adam_step(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, scratch[1:2])
        assert torch.allclose(p, ref.detach(), atol=1e-6, rtol=1e-5)


def test_adam_device_step_counter_matches_torch(cuda):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(11)
    n = 50_001
    p0 = torch.randn(n, generator=g).cuda()
    p = p0.clone()
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.zeros(3, device="cuda")
    for _ in range(4):
        grad = torch.randn(n, generator=g).cuda()
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step_dev(p, grad, m, v, 2e-3, 0.9, 0.999, 1e-8, 0.0, state)
        assert torch.allclose(p, ref.detach(), atol=1e-6, rtol=1e-5)
    assert state[0].item() == 4
