"""GPU: the three epilogue variants of conv_fwd_kernel (0 = row stores from registers, 1 = smem transpose + coalesced
stores, 2 = swizzled smem slab + TMA tile store) must produce identical 16-bit outputs (same accumulators, same math)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture
def opts():
    from gdl_b200 import ops
    ops.set_option("conv_rows", 0)  # keep every case on conv_fwd_kernel
    yield ops.set_option
    ops.set_option("conv_rows", 1)
    ops.set_option("conv_epilogue", 2)


@pytest.mark.parametrize("n,h,w,chans,cout,k,dtype,kw", [
    (1, 1, 5000, [64], 256, 1, torch.bfloat16, {}),                      # flat GEMM, ragged last tile, 4 slabs of 64
    (2, 24, 256, [64], 448, 3, torch.bfloat16, {"relu": True}),          # 2 n-tiles of 224 -> slabs of 32
    (2, 20, 36, [128], 256, 3, torch.bfloat16, {"bias": True}),          # 2-D pixel tile (TH x TW box), ragged both ways
    (1, 16, 128, [64, 64], 80, 3, torch.bfloat16, {"bias": True, "relu": True}),   # BN = 80 -> slabs of 16
    (3, 9, 130, [32], 32, 3, torch.float16, {"gelu": True, "bias": True}),
    (1, 1, 777, [256], 1024, 1, torch.float16, {"oscale": True, "bias": True}),
    (2, 8, 128, [64], 64, 3, torch.bfloat16, {"bias": True, "relu": True}),        # halo mode of conv_fwd_kernel
    (1, 1, 3000, [64], 64, 1, torch.bfloat16, {"bias": True, "out32": True, "res": "f32"}),   # fp32 residual stream
    (2, 17, 33, [128], 320, 1, torch.bfloat16, {"out32": True, "res": "f32", "oscale": True}),  # 2 n-tiles of 160: fp32 slabs of 32
    (1, 12, 128, [64], 20, 3, torch.bfloat16, {"bias": True, "out32": True}),                 # fp32 logits, Cout = 20
    (2, 10, 96, [256], 128, 3, torch.float16, {"res": "16", "relu": True}),                   # 16-bit residual (dgrad + skip)
])
def test_epilogue_variants_identical(cuda, opts, n, h, w, chans, cout, k, dtype, kw):
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(h * w + cout)
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(dtype).cuda() for c in chans]
    ctot = sum(chans)
    wt = (torch.randn(cout, ctot, k, k, generator=g) / (k * k * ctot) ** 0.5).cuda()
    wp = ops.pack_conv_weight(wt, dtype)
    args = dict(relu=kw.get("relu", False), gelu=kw.get("gelu", False))
    if kw.get("bias"):
        args["bias"] = torch.randn(cout, generator=g).cuda()
    if kw.get("oscale"):
        args["oscale"] = torch.randn(cout, generator=g).cuda()
    res = None
    if kw.get("res"):
        res = torch.randn(n, h, w, cout, generator=g).cuda()
        res = res if kw["res"] == "f32" else res.to(dtype)
        args["residual"] = res
    if kw.get("out32"):
        args["out_dtype"] = torch.float32
    outs = {}
    for mode in (0, 1, 2):
        opts("conv_epilogue", mode)
        outs[mode] = ops.conv2d_fwd(srcs, wp, cout, k, k, k // 2, k // 2, **args)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    x = torch.cat([t.float() for t in srcs], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(x, wp.view(cout, k, k, ctot).float().permute(0, 3, 1, 2), args.get("bias"), padding=k // 2)
    if "oscale" in args:
        ref = ref * args["oscale"].view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.float().permute(0, 3, 1, 2)
    ref = F.relu(ref) if args["relu"] else (F.gelu(ref) if args["gelu"] else ref)
    err = ((outs[2].float() - ref.permute(0, 2, 3, 1)).abs().max() / ref.abs().max()).item()
    assert err < (1e-4 if kw.get("out32") else 6e-3)


def test_tma_store_epilogue_respects_channel_slices(cuda, opts):
    """out is a channel slice of a wider buffer (ldo > Cout): neighbours must stay untouched."""
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(2, 10, 128, 64, generator=g) * 0.5).bfloat16().cuda()
    wp = ops.pack_conv_weight((torch.randn(128, 64, 1, 1, generator=g) / 8).cuda(), torch.bfloat16)
    res = {}
    for mode in (0, 2):
        opts("conv_epilogue", mode)
        out = torch.full((2, 10, 128, 192), 7.0, dtype=torch.bfloat16, device="cuda")
        ops.conv2d_fwd([x], wp, 128, 1, 1, 0, 0, out=out[..., 32:160])
        assert (out[..., :32] == 7).all() and (out[..., 160:] == 7).all()
        res[mode] = out
    assert torch.equal(res[0], res[2])
