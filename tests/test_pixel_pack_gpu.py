"""GPU: narrow (16/32-channel) 3x3 convs run in their pixel-packed form (f adjacent pixels = one 64-channel pixel,
block-Toeplitz weights) must equal the plain path and fp32 torch, forward, dgrad and wgrad."""
import pytest
import torch
import torch.nn.functional as F

import cpu_kernel_emulation as emu

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("co,ci,f", [(16, 16, 4), (16, 32, 2), (32, 16, 4), (32, 32, 2), (16, 16, 2)])
def test_widen_and_fold_kernels_match_their_definition(cuda, co, ci, f):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(co * 100 + ci + f)
    w = torch.randn(co, ci, 3, 3, generator=g).cuda()
    for mode in (0, 1):
        wp = ops.pack_conv_weight(w, BF, mode)
        a, b = (co, ci) if mode == 0 else (ci, co)
        wide = ops.widen_conv_weight(wp, a, b, 3, f)
        assert torch.equal(wide.cpu(), emu.widen_conv_weight(wp.cpu(), a, b, 3, f))
    dw = torch.randn(f * co, 9 * f * ci, generator=g).cuda()
    out = torch.empty(co, ci, 3, 3, device="cuda")
    ops.fold_widened_wgrad(dw, out, f)
    want = emu.fold_widened_wgrad(dw.cpu(), torch.empty(co, ci, 3, 3), f)
    assert torch.allclose(out.cpu(), want, atol=1e-5)
    ops.fold_widened_wgrad(dw, out, f, accumulate=True)
    assert torch.allclose(out.cpu(), 2 * want, atol=2e-5)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 6, 128, 16, 16), (1, 5, 256, 32, 16), (2, 4, 64, 16, 32), (1, 3, 512, 32, 32),
                                          (1, 7, 12, 16, 16)])
def test_engine_pixel_packed_conv_equals_plain_and_fp32(cuda, n, h, w, ci, co):
    from gdl_b200 import ops
    from gdl_b200.engine import Act, Engine
    g = torch.Generator().manual_seed(w + ci + co)
    x = (torch.randn(n, h, w, ci, generator=g) * 0.5).to(BF).cuda()
    wt = torch.nn.Parameter((torch.randn(co, ci, 3, 3, generator=g) / (9 * ci) ** 0.5).cuda())
    dy = (torch.randn(n, h, w, co, generator=g) * 0.5).to(BF).cuda()
    res = {}
    try:
        for mode in (0, 1):
            ops.set_option("pixel_pack", mode)
            eng = Engine(BF, training=True)
            a = Act(x, needs_grad=True)
            rc = eng.conv_raw([a], wt, 1, 1, out_dtype=torch.float32)
            assert rc.pixel_packed == bool(mode)
            eng.conv_backward(rc, dy)
            res[mode] = (rc.x, a.gsrcs[0][0].float(), eng.param_grads[id(wt)].clone())
    finally:
        ops.set_option("pixel_pack", 1)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().to(BF).float().requires_grad_(True)
    y = F.conv2d(xr, wr, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = (y.permute(0, 2, 3, 1), xr.grad.permute(0, 2, 3, 1), wr.grad)
    for name, plain, packed, r in zip(("y", "dx", "dw"), res[0], res[1], ref):
        assert packed.shape == r.shape
        assert _relerr(plain, r) < 4e-3, name   # dx is rounded to bf16 by the dgrad epilogue
        assert _relerr(packed, r) < 4e-3, name
        assert _relerr(packed, plain) < 4e-3, name
    assert _relerr(res[1][0], ref[0]) < 2e-5 and _relerr(res[1][2], ref[2]) < 2e-5  # fp32 outputs: accumulation order only


def test_pixel_packed_head_conv_with_bias_and_padded_dlogits(cuda):
    """The segmentation head (16 -> K=5 logits, bias, fp32 out) and its backward on 16-channel zero-padded dlogits."""
    from gdl_b200 import ops
    from gdl_b200.engine import Act, Engine
    g = torch.Generator().manual_seed(7)
    n, h, w, ci, k = 2, 5, 64, 16, 5
    x = (torch.randn(n, h, w, ci, generator=g) * 0.5).to(BF).cuda()
    wt = torch.nn.Parameter((torch.randn(k, ci, 3, 3, generator=g) / 12).cuda())
    bias = torch.nn.Parameter(torch.randn(k, generator=g).cuda())
    d16 = torch.zeros(n, h, w, 16, dtype=BF).cuda()
    d16[..., :k] = (torch.randn(n, h, w, k, generator=g) * 0.5).to(BF).cuda()
    res = {}
    try:
        for mode in (0, 1):
            ops.set_option("pixel_pack", mode)
            eng = Engine(BF, training=True)
            a = Act(x, needs_grad=True)
            logits = eng.conv_head(a, wt, bias, 1)
            assert logits.shape == (n, h, w, k) and logits.dtype == torch.float32
            eng.head_backward(d16)
            res[mode] = (logits, a.gsrcs[0][0].float(), eng.param_grads[id(wt)].clone(), eng.param_grads[id(bias)].clone())
    finally:
        ops.set_option("pixel_pack", 1)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().to(BF).float().requires_grad_(True)
    br = bias.detach().clone().requires_grad_(True)
    y = F.conv2d(xr, wr, br, padding=1)
    y.backward(d16[..., :k].float().permute(0, 3, 1, 2))
    ref = (y.permute(0, 2, 3, 1), xr.grad.permute(0, 2, 3, 1), wr.grad, br.grad)
    for name, plain, packed, r in zip(("logits", "dx", "dw", "db"), res[0], res[1], ref):
        assert packed.shape == r.shape, name
        assert _relerr(plain, r) < 4e-3 and _relerr(packed, r) < 4e-3, name
