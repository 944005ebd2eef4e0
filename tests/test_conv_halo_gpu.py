"""GPU: the halo-tile paths of the tensor-core convs (3 horizontal taps served by one TMA row tile through
row-shifted UMMA descriptors) must give the same results as the per-tap paths and as fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.fixture
def halo_switch():
    from gdl_b200 import ops
    yield ops.set_option
    ops.set_option("conv_halo", 1)
    ops.set_option("wgrad_halo", 1)


@pytest.mark.parametrize("n,h,w,chans,cout", [
    (2, 4, 128, [64], 64), (1, 6, 256, [64, 128], 32), (2, 3, 384, [256, 64, 64], 128), (1, 2, 128, [64], 16),
    (1, 5, 512, [128], 64), (1, 4, 256, [16], 16), (2, 3, 128, [32], 16), (1, 3, 128, [32, 32], 32), (1, 2, 128, [16], 5),
])
def test_fwd_halo_equals_per_tap_and_fp32(cuda, halo_switch, n, h, w, chans, cout):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(w + cout)
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(BF).cuda() for c in chans]
    ctot = sum(chans)
    wt = (torch.randn(cout, ctot, 3, 3, generator=g) / (9 * ctot) ** 0.5).cuda()
    wp = ops.pack_conv_weight(wt, BF)
    outs = {}
    for mode in (0, 1):
        halo_switch("conv_halo", mode)
        outs[mode] = ops.conv2d_fwd(srcs, wp, cout, 3, 3, 1, 1, out_dtype=torch.float32)
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(x, wp.view(cout, 3, 3, ctot).float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    assert _relerr(outs[0], ref) < 2e-3
    assert _relerr(outs[1], ref) < 2e-3
    assert _relerr(outs[1], outs[0]) < 1e-5  # same products, same accumulation order per tap group


@pytest.mark.parametrize("n,h,w,chans,cout", [
    (2, 4, 64, [64], 64), (1, 6, 128, [64, 128], 32), (2, 3, 256, [256, 64, 64], 128), (1, 2, 128, [64], 16),
    (1, 5, 192, [128], 256), (1, 4, 256, [16], 16), (2, 3, 128, [32], 16), (1, 3, 64, [32, 32], 32),
])
def test_wgrad_halo_equals_per_tap_and_autograd(cuda, halo_switch, n, h, w, chans, cout):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(w + cout + 1)
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(BF).cuda() for c in chans]
    ctot = sum(chans)
    dy = (torch.randn(n, h, w, cout, generator=g) * 0.5).to(BF).cuda()
    outs = {}
    for mode in (0, 1):
        halo_switch("wgrad_halo", mode)
        dw = torch.zeros(cout, 9 * ctot, device="cuda")
        ops.conv2d_wgrad(srcs, dy, 3, 3, 1, 1, dw)
        outs[mode] = dw
    x = torch.cat([t.double() for t in srcs], 3).permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.grad.conv2d_weight(x, (cout, ctot, 3, 3), dy.double().permute(0, 3, 1, 2).contiguous(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(cout, 9 * ctot)
    assert _relerr(outs[0], ref) < 2e-3
    assert _relerr(outs[1], ref) < 2e-3
