"""GPU: the weight-stationary row-rolling 3x3 kernel (csrc/conv3x3_rows.cu, Cout <= 64) against the per-tile kernel
and fp32 torch on the same bf16/fp16 operands — forward, epilogue options, ragged widths, every group height G."""
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
    ops.set_option("conv_rows", 1)


@pytest.mark.parametrize("n,h,w,chans,cout,dtype", [
    (2, 8, 128, [64], 64, torch.bfloat16),            # G = 4
    (1, 6, 256, [64, 128], 32, torch.bfloat16),       # G = 2 (6 % 4 != 0), two sources, two column tiles
    (2, 16, 200, [256, 64, 64], 16, torch.bfloat16),  # G = 8, ragged width, 6 k-chunks
    (1, 4, 64, [64], 48, torch.bfloat16),             # narrow image, BN = 48 -> G = 4
    (3, 12, 384, [128], 64, torch.float16),           # fp16 operands, 3 column tiles
    (1, 32, 128, [64, 64, 64, 64, 64], 32, torch.bfloat16),
    (2, 8, 128, [64], 5, torch.bfloat16),             # Cout not a multiple of 16 (logits)
])
def test_rows_kernel_equals_tile_kernel_and_fp32(cuda, rows_switch, n, h, w, chans, cout, dtype):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(h * w + cout)
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(dtype).cuda() for c in chans]
    ctot = sum(chans)
    wt = (torch.randn(cout, ctot, 3, 3, generator=g) / (9 * ctot) ** 0.5).cuda()
    wp = ops.pack_conv_weight(wt, dtype)
    outs = {}
    for mode in (0, 1):
        rows_switch("conv_rows", mode)
        outs[mode] = ops.conv2d_fwd(srcs, wp, cout, 3, 3, 1, 1, out_dtype=torch.float32)
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(x, wp.view(cout, 3, 3, ctot).float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    assert _relerr(outs[0], ref) < 2e-3
    assert _relerr(outs[1], ref) < 2e-3
    assert _relerr(outs[1], outs[0]) < 1e-4  # same products, fp32 accumulation order differs
    # 16-bit outputs leave through the smem-staged TMA store (ragged widths are clipped by the tensor map)
    if cout % 8 == 0:
        y16 = ops.conv2d_fwd(srcs, wp, cout, 3, 3, 1, 1, relu=True)
        assert y16.dtype == dtype and torch.equal(y16, F.relu(outs[1]).to(dtype))


def test_rows_kernel_epilogue_options(cuda, rows_switch):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(11)
    n, h, w, c, cout = 2, 8, 256, 64, 64
    x = (torch.randn(n, h, w, c, generator=g) * 0.5).bfloat16().cuda()
    wt = (torch.randn(cout, c, 3, 3, generator=g) / 24).cuda()
    wp = ops.pack_conv_weight(wt, torch.bfloat16)
    bias = torch.randn(cout, generator=g).cuda()
    gamma = torch.randn(cout, generator=g).cuda()
    res32 = torch.randn(n, h, w, cout, generator=g).cuda()
    res16 = res32.bfloat16()
    acc = F.conv2d(x.float().permute(0, 3, 1, 2), wp.view(cout, 3, 3, c).float().permute(0, 3, 1, 2),
                   padding=1).permute(0, 2, 3, 1)
    cases = [
        (dict(bias=bias, relu=True), torch.bfloat16, F.relu(acc + bias)),
        (dict(bias=bias, gelu=True), torch.float32, F.gelu(acc + bias)),
        (dict(bias=bias, oscale=gamma, residual=res32), torch.float32, (acc + bias) * gamma + res32),
        (dict(residual=res16), torch.bfloat16, acc + res16.float()),
    ]
    for kw, odt, ref in cases:
        for mode in (0, 1):
            rows_switch("conv_rows", mode)
            y = ops.conv2d_fwd([x], wp, cout, 3, 3, 1, 1, out_dtype=odt, **kw)
            assert _relerr(y, ref) < (1e-4 if odt == torch.float32 else 6e-3), (sorted(kw), mode)
    # a channel slice of a wider output buffer (ldo > Cout) and a source that is a channel slice (ld > C)
    rows_switch("conv_rows", 1)
    wide_in = (torch.randn(n, h, w, 128, generator=g) * 0.5).bfloat16().cuda()
    out = torch.zeros(n, h, w, 96, dtype=torch.bfloat16, device="cuda")
    ops.conv2d_fwd([wide_in[..., 64:]], wp, cout, 3, 3, 1, 1, out=out[..., 16:80])
    ref = F.conv2d(wide_in[..., 64:].float().permute(0, 3, 1, 2), wp.view(cout, 3, 3, c).float().permute(0, 3, 1, 2),
                   padding=1).permute(0, 2, 3, 1)
    assert _relerr(out[..., 16:80], ref) < 6e-3
    assert out[..., :16].abs().max() == 0 and out[..., 80:].abs().max() == 0
