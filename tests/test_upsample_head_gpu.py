"""GPU: the fused head — bilinear upsample (align_corners=False) + loss / argmax without the (N,K,H,W) logits.

Reference arithmetic: F.interpolate(logits, size=image.shape[2:], mode="bilinear", align_corners=False)
(models/segmentation/segformer.py:47-57, models/segmentation/dofa.py:90-105) followed by the loss of training_step or by
softmax.argmax of the eval steps (tasks_with_models/segmentation_segformer.py:218-243,268-271).
Bars: loss to 1e-6 against F.interpolate + F.cross_entropy in fp32, the low-resolution gradient to 1e-5 (relative) against
torch autograd, class maps torch.equal to the unfused kernels (same interpolation arithmetic, operation for operation).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (N, h, w, H, W, K): SegFormer's x4 head, DOFA's 144 -> 512 and 18 -> 512 (non-integer scales), ragged sizes, downscale
GEOMS = [(2, 16, 16, 64, 64, 5), (1, 18, 18, 64, 64, 5), (3, 5, 7, 64, 96, 3), (1, 3, 3, 96, 96, 19), (2, 40, 24, 20, 12, 4)]


def _data(n, h, w, hh, ww, k, seed=0, ignore=False, u8=False):
    g = torch.Generator().manual_seed(seed + h * w + k)
    lr = (torch.randn(n, h, w, k, generator=g) * 2).cuda()
    t = torch.randint(0, max(k, 2), (n, hh, ww), generator=g)
    if ignore:
        t[torch.rand(n, hh, ww, generator=g) < 0.1] = 255 if u8 else -100
    t = t.to(torch.uint8).cuda() if u8 else t.cuda()
    return lr, t


def _torch_logits(lr, hh, ww):
    return F.interpolate(lr.permute(0, 3, 1, 2), size=(hh, ww), mode="bilinear", align_corners=False)


@pytest.mark.parametrize("n,h,w,hh,ww,k", GEOMS)
def test_fused_ce_matches_interpolate_plus_cross_entropy(cuda, n, h, w, hh, ww, k):
    from gdl_b200 import ops
    lr, t = _data(n, h, w, hh, ww, k)
    spec = ops.LossSpec(1.0, 0.0, ignore_index=-100)
    coeff, stats = ops.upsample_ce_fwd(lr, t, spec)
    lr_ref = lr.clone().requires_grad_(True)
    ref = F.cross_entropy(_torch_logits(lr_ref, hh, ww), t)
    ref.backward()
    assert abs(coeff[0].item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item())) + 2e-6
    d = torch.zeros_like(lr)
    ops.upsample_ce_bwd(lr, t, spec, coeff, None, d)
    err = (d - lr_ref.grad).norm() / lr_ref.grad.norm()
    print(f"{h}x{w}->{hh}x{ww} K={k}: loss {coeff[0].item():.7f} vs {ref.item():.7f}, grad rel err {err.item():.2e}")
    assert err < 1e-5
    # the unfused kernels (materialised logits) give the same statistics bit for bit, and the same class map
    up = ops.bilinear_fwd(lr, hh, ww)
    coeff_u, stats_u = ops.seg_loss_fwd(up, t, spec)
    assert torch.equal(stats, stats_u) and torch.equal(coeff, coeff_u)
    assert torch.equal(ops.upsample_argmax(lr, hh, ww), ops.argmax_classes(up))
    agree = (ops.upsample_argmax(lr, hh, ww) == _torch_logits(lr, hh, ww).argmax(1)).float().mean().item()
    assert agree > 0.9999  # fp32 re-association can flip exact near-ties only


@pytest.mark.parametrize("n,h,w,hh,ww,k", [(16, 128, 128, 512, 512, 5), (8, 144, 144, 512, 512, 5), (8, 18, 18, 512, 512, 5)])
def test_fused_ce_at_baseline_shapes(cuda, n, h, w, hh, ww, k):
    """BASELINE configs[2] (SegFormer-B2, B = 16: 128^2 -> 512^2) and configs[3] (DOFA: 144^2 and 18^2 -> 512^2)"""
    test_fused_ce_matches_interpolate_plus_cross_entropy(cuda, n, h, w, hh, ww, k)


@pytest.mark.parametrize("u8", [False, True])
def test_fused_dice_ce_label_smoothing_ignore_index(cuda, u8):
    from gdl_b200 import ops
    n, h, w, hh, ww, k = 2, 12, 12, 48, 48, 5
    lr, t = _data(n, h, w, hh, ww, k, seed=3, ignore=True, u8=u8)
    spec = ops.LossSpec(0.7, 0.5, label_smoothing=0.1, ignore_index=255 if u8 else -100)
    coeff, _ = ops.upsample_ce_fwd(lr, t, spec)
    up = ops.bilinear_fwd(lr, hh, ww)
    coeff_u, _ = ops.seg_loss_fwd(up, t, spec)
    assert torch.equal(coeff, coeff_u)
    d = torch.zeros(n, h, w, 16, dtype=torch.bfloat16, device="cuda")  # the padded 16-bit operand of the head's backward
    gscale = torch.full((1,), 4.0, device="cuda")
    ops.upsample_ce_bwd(lr, t, spec, coeff, gscale, d)
    du = torch.empty_like(up)
    ops.seg_loss_bwd(up, t, spec, coeff_u, gscale, du)
    ref = ops.bilinear_bwd(du, h, w)
    assert (d[..., k:] == 0).all()
    err = (d[..., :k].float() - ref).norm() / ref.norm()
    print("dice+ce fused vs unfused low-res gradient:", err.item())
    assert err < 5e-3  # one bf16 rounding of the output


def test_fused_binary_head(cuda):
    from gdl_b200 import ops
    n, h, w, hh, ww = 2, 16, 16, 64, 64
    lr, t = _data(n, h, w, hh, ww, 1, seed=5)
    spec = ops.LossSpec(1.0, 1.0, ignore_index=None)
    coeff, stats = ops.upsample_ce_fwd(lr, t, spec)
    up = ops.bilinear_fwd(lr, hh, ww)
    coeff_u, stats_u = ops.seg_loss_fwd(up, t, spec)
    assert torch.equal(coeff, coeff_u) and torch.equal(stats, stats_u)
    d = torch.zeros_like(lr)
    ops.upsample_ce_bwd(lr, t, spec, coeff, None, d)
    du = torch.empty_like(up)
    ops.seg_loss_bwd(up, t, spec, coeff_u, None, du)
    ref = ops.bilinear_bwd(du, h, w)
    assert (d - ref).norm() / ref.norm() < 1e-5
    assert torch.equal(ops.upsample_argmax(lr, hh, ww, 0.5), ops.argmax_classes(up, 0.5))


def test_fused_head_is_reproducible_and_rejects_bad_arguments(cuda):
    from gdl_b200 import ops
    lr, t = _data(2, 16, 16, 64, 64, 5, seed=7)
    spec = ops.LossSpec(1.0, 0.0, ignore_index=-100)
    a = ops.upsample_ce_fwd(lr, t, spec)
    b = ops.upsample_ce_fwd(lr, t, spec)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    d1, d2 = torch.zeros_like(lr), torch.zeros_like(lr)
    ops.upsample_ce_bwd(lr, t, spec, a[0], None, d1)
    ops.upsample_ce_bwd(lr, t, spec, a[0], None, d2)
    assert torch.equal(d1, d2)
    with pytest.raises(NotImplementedError):
        ops.upsample_ce_fwd(torch.zeros(1, 4, 4, 40, device="cuda"), torch.zeros(1, 8, 8, dtype=torch.int64, device="cuda"), spec)


def test_segformer_fused_head_step_equals_unfused_step(cuda):
    """FusedTrainer on SegFormer: fused head (default) vs the materialised logits — identical loss (same statistics pass
    arithmetic), gradients equal up to the summation order of the low-resolution logit gradient; predict_classes equals
    softmax.argmax of forward()."""
    from gdl_b200 import ops
    from gdl_b200.models.segformer import SegFormer
    from gdl_b200.trainer import FusedTrainer
    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (4, 4, 4), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    raw = (t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 128, 128, 3), generator=g).cuda()).to(torch.uint8)
    res = {}
    old = ops.option("fused_head")
    try:
        for fused in (1, 0):
            ops.set_option("fused_head", fused)
            torch.manual_seed(3)
            m = SegFormer("mit_b0", in_channels=3, num_classes=5).cuda().train()
            tr = FusedTrainer(m, ops.LossSpec(1.0, 0.3, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.2] * 3)
            loss = tr.forward_backward(raw, t)
            res[fused] = (loss.item(), tr.gflat.clone())
    finally:
        ops.set_option("fused_head", old)
    assert res[1][0] == res[0][0]
    rel = ((res[1][1] - res[0][1]).norm() / res[0][1].norm()).item()
    print("fused vs unfused head: flat gradient rel diff", rel)
    assert rel < 2e-3
    m.eval()
    x = torch.randn(2, 3, 128, 128, generator=g).cuda()
    with torch.no_grad():
        ref = m(x).softmax(dim=1).argmax(dim=1)
    assert torch.equal(m.predict_classes(x), ref)
