"""GPU: the fp32-accurate mode of the tensor-core convolutions (two bf16 planes per operand, ops.conv2d_*_bf16x2) meets the tolerance
north_star states for fp32 — 1e-5, max norm — against the float64 convolution of the same fp32 operands (torch's own fp32
convolution, TF32 off, is reported beside it): forward, data gradient and weight gradient, implicit-GEMM / row-streaming / pointwise kernels, virtual concat.  (The 16-bit training path
is compared with the autocast reference instead: one bf16 rounding is 2^-9.)"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # (N, H, W, [source channels], Cout, R)
    (2, 32, 32, [64], 64, 3),
    (2, 32, 48, [128, 64], 256, 3),
    (1, 64, 128, [64], 32, 3),
    (4, 16, 16, [256], 128, 1),
    (2, 24, 24, [512, 256], 256, 3),
]


def _err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.parametrize("n,h,w,cins,cout,r", CASES)
def test_bf16x2_convolution_meets_1e5(cuda, n, h, w, cins, cout, r):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(h * w + cout)
    srcs = [torch.randn(n, h, w, c, generator=g).cuda() for c in cins]
    ctot, pad = sum(cins), r // 2
    wt = (torch.randn(cout, ctot, r, r, generator=g) / (ctot * r * r) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    dy = torch.randn(n, h, w, cout, generator=g).cuda()
    # the oracle is the convolution in float64 (exact to ~1e-16); torch's own fp32 convolution is measured against it too —
    # over the 2 000 .. 8 000-term contractions of a weight gradient fp32 accumulation alone costs either implementation
    # several 1e-6
    x = torch.cat(srcs, dim=3).permute(0, 3, 1, 2).contiguous().double().requires_grad_(True)
    wr = wt.double().clone().requires_grad_(True)
    ref = F.conv2d(x, wr, bias.double(), padding=pad)
    ref.backward(dy.double().permute(0, 3, 1, 2))
    x32 = x.detach().float().requires_grad_(True)
    w32 = wt.clone().requires_grad_(True)
    y32 = F.conv2d(x32, w32, bias, padding=pad)
    y32.backward(dy.permute(0, 3, 1, 2))
    t_f, t_d, t_w = _err(y32, ref), _err(x32.grad, x.grad), _err(w32.grad, wr.grad)
    # forward
    y = ops.conv2d_fwd_bf16x2(srcs, wt, pad, pad, bias=bias)
    e_f = _err(y.permute(0, 3, 1, 2), ref)
    # one bf16 product for comparison
    wp = ops.pack_conv_weight(wt, torch.bfloat16)
    y16 = ops.conv2d_fwd([s.to(torch.bfloat16) for s in srcs], wp, cout, r, r, pad, pad, out_dtype=torch.float32, bias=bias)
    e_16 = _err(y16.permute(0, 3, 1, 2), ref)
    # data gradient
    dx = ops.conv2d_fwd_bf16x2([dy], wt, r - 1 - pad, r - 1 - pad, mode=1)
    e_d = _err(dx.permute(0, 3, 1, 2), x.grad)
    # weight gradient
    dw = ops.conv2d_wgrad_bf16x2(srcs, dy, r, r, pad, pad)
    e_w = _err(dw.view(cout, r, r, ctot).permute(0, 3, 1, 2), wr.grad)
    print(f"{cins}->{cout} k{r} @{h}x{w}: bf16x2 fwd {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e}   (torch fp32 conv: {t_f:.2e} {t_d:.2e} "
          f"{t_w:.2e}; one bf16 product: {e_16:.2e})")
    assert e_f < 1e-5 and e_d < 1e-5 and e_w < 1e-5
    assert e_16 > 20 * e_f  # the split really buys the accuracy
