"""GPU: the paired-tap row-streaming weight-gradient kernel (csrc/wgrad3x3_rows.cu, Cout = 16/32/64) against the
generic wgrad kernel and fp32 torch on the same 16-bit operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.fixture
def rows_switch():
    from gdl_b200 import ops
    yield ops.set_option
    ops.set_option("wgrad_rows", 1)


@pytest.mark.parametrize("n,h,w,chans,cout,dtype", [
    (2, 8, 128, [64], 64, torch.bfloat16),
    (1, 37, 200, [256, 64], 32, torch.bfloat16),      # ragged row blocks and ragged width, 5 slabs
    (3, 16, 256, [128], 16, torch.bfloat16),
    (2, 64, 128, [64, 64, 64], 64, torch.float16),    # several row blocks per column
    (1, 4, 64, [64], 64, torch.bfloat16),             # smallest supported image
    (2, 12, 384, [64, 128, 64], 32, torch.bfloat16),
])
def test_wgrad_rows_equals_generic_and_fp32(cuda, rows_switch, n, h, w, chans, cout, dtype):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(h * w + cout)
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(dtype).cuda() for c in chans]
    dy = (torch.randn(n, h, w, cout, generator=g) * 0.5).to(dtype).cuda()
    ctot = sum(chans)
    prefill = torch.randn(cout, 9 * ctot, generator=g).cuda()
    outs = {}
    for mode in (0, 1):
        rows_switch("wgrad_rows", mode)
        dw = prefill.clone()
        ops.conv2d_wgrad(srcs, dy, 3, 3, 1, 1, dw)  # accumulates into dw
        outs[mode] = dw - prefill
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    wref = torch.zeros(cout, ctot, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(x, wref, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ref = wref.grad.permute(0, 2, 3, 1).reshape(cout, 9 * ctot)  # [cout][(ky,kx,c)]
    scale = ref.abs().max()
    assert (outs[0] - ref).abs().max() / scale < 2e-3
    assert (outs[1] - ref).abs().max() / scale < 2e-3
    assert (outs[1] - outs[0]).abs().max() / scale < 1e-4
