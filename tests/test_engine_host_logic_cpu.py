"""CPU: host logic of the engine / model graph / trainer, with the CUDA kernels replaced by the fp32
torch emulation in tests/cpu_kernel_emulation.py.  With fp32 "kernels" the engine's hand-written
backward must reproduce the oracle's autograd to float accuracy — this pins the graph wiring
(virtual concat order, gradient-source bookkeeping, residual / downsample branches, im2col routing)."""
import pytest
import torch
import torch.nn.functional as F

import cpu_kernel_emulation as emu


def _pair(enc, cin, k):
    from gdl_b200.models.unetpp import UnetPlusPlus
    from oracle.unetpp import UnetPlusPlusOracle
    torch.manual_seed(0)
    ora = UnetPlusPlusOracle(enc, cin, k)
    with torch.no_grad():
        for m in ora.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    prod = UnetPlusPlus(enc, in_channels=cin, classes=k, compute_dtype=torch.float32)
    prod.load_state_dict(ora.state_dict())
    return ora, prod


@pytest.fixture
def f64(monkeypatch):
    """float64 everywhere: in fp32 a single pre-activation within 1e-5 of zero flips its ReLU mask
    between two summation orders and moves a BN-bias gradient by percents (observed), which would
    hide real wiring bugs behind a loose tolerance."""
    emu.set_work_dtype(torch.float64)
    yield torch.float64
    emu.set_work_dtype(torch.float32)


@pytest.mark.parametrize("enc,cin,k", [("resnet18", 3, 5), ("resnet50", 4, 3), ("resnext50_32x4d", 3, 2)])
def test_engine_backward_equals_oracle_autograd(monkeypatch, f64, enc, cin, k):
    from gdl_b200.engine import Act, Engine
    emu.install(monkeypatch)
    ora, prod = _pair(enc, cin, k)
    ora, prod = ora.double(), prod.double()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, 64, 64, generator=g).double()
    t = torch.randint(0, k, (2, 64, 64), generator=g)
    ora.train()
    ref = ora(x)
    F.cross_entropy(ref, t).backward()

    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    xin = emu.normalize_to_nhwc(x, True, torch.float64, (cin + 7) // 8 * 8)
    with torch.no_grad():  # the product runs inside autograd.Function.forward / FusedTrainer (no autograd)
        logits = prod.run(eng, Act(xin, needs_grad=False))
    assert torch.allclose(logits.permute(0, 3, 1, 2), ref, atol=1e-9, rtol=1e-9)
    d = logits.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.cross_entropy(d, t).backward()
    d16 = torch.zeros(2, 64, 64, 16, dtype=torch.float64)
    d16[..., :k] = d.grad.permute(0, 2, 3, 1)
    with torch.no_grad():
        eng.head_backward(d16)
        eng.backward()
    refg = dict(ora.named_parameters())
    for n, p in prod.named_parameters():
        got = eng.param_grads[id(p)]
        want = refg[n].grad
        err = (got - want).norm() / (want.norm() + 1e-12)
        assert err < 1e-8, f"{n}: {err}"
    # running statistics were updated exactly once
    for (n, a), (_, b) in zip(prod.named_buffers(), ora.named_buffers()):
        if "running" in n:
            assert torch.allclose(a, b, atol=1e-9, rtol=1e-9), n
    assert not eng.tape


def test_trainer_flat_buffers_and_step(monkeypatch):
    from gdl_b200.trainer import FusedTrainer
    emu.install(monkeypatch)
    _, prod = _pair("resnet18", 3, 4)
    prod.train()
    tr = FusedTrainer(prod, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-2, mean=[0.5] * 3, std=[0.25] * 3,
                      clip_grad_norm=1.0)
    assert tr.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in prod.parameters())  # 16-byte aligned slots
    # parameters are views of the flat buffer, gradients views of the flat gradient buffer
    p0 = next(prod.parameters())
    assert p0.data_ptr() == tr.flat.data_ptr() and p0.grad.data_ptr() == tr.gflat.data_ptr()
    g = torch.Generator().manual_seed(3)
    t = torch.randint(0, 4, (2, 2, 2), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    raw = (t.unsqueeze(-1) * 60 + torch.randint(0, 20, (2, 32, 32, 3), generator=g)).to(torch.uint8)
    losses = [float(tr.step(raw, t)) for _ in range(6)]
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]
    # state_dict still exposes the (updated) parameters under the reference's key names
    sd = prod.state_dict()
    assert "decoder.blocks.x_0_0.conv1.0.weight" in sd and sd["encoder.conv1.weight"].data_ptr() == p0.data_ptr()


def test_trainer_flat_buffer_slots_are_16_byte_aligned(monkeypatch):
    """Parameters whose element count is not a multiple of 4 (a 5-class bias) must not misalign what follows them: the
    wgrad kernel accumulates into the flat gradient buffer with 128-bit reductions."""
    from gdl_b200.trainer import FusedTrainer
    emu.install(monkeypatch)
    _, prod = _pair("resnet18", 3, 5)
    extra = torch.nn.Parameter(torch.ones(7))
    prod.register_parameter("odd_sized_first", extra)
    prod._parameters.move_to_end("odd_sized_first", last=False) if hasattr(prod._parameters, "move_to_end") else None
    tr = FusedTrainer(prod.train(), emu.LossSpec(1.0, 0.0, ignore_index=-100), mean=[0.5] * 3, std=[0.25] * 3)
    base = tr.flat.data_ptr()
    for p in tr.params:
        assert (p.data_ptr() - base) % 16 == 0 and (p.grad.data_ptr() - tr.gflat.data_ptr()) % 16 == 0
    assert torch.equal(extra.detach(), torch.ones(7))  # values survive the move into the flat buffer
    used = sum(p.numel() for p in tr.params)
    covered = torch.zeros(tr.flat.numel(), dtype=torch.bool)
    for p in tr.params:
        off = (p.data_ptr() - base) // tr.flat.element_size()
        covered[off:off + p.numel()] = True
    assert tr.flat.numel() > used and int(covered.sum()) == used and not tr.flat[~covered].any()  # gaps are zero


def test_trainer_loss_scaling(monkeypatch, f64):
    """fp16 loss scaling of the fused trainer (what Lightning's GradScaler does around the reference's backward): a static
    scale leaves the update unchanged (exact powers of two), the dynamic rule skips a step with a non-finite gradient and
    halves the scale, and doubles it after `growth_interval` clean steps."""
    from gdl_b200.trainer import FusedTrainer
    emu.install(monkeypatch)
    g = torch.Generator().manual_seed(0)
    raw = torch.randint(0, 256, (2, 32, 32, 3), generator=g, dtype=torch.uint8)
    t = torch.randint(0, 4, (2, 32, 32), generator=g)

    def trainer(**kw):
        _, prod = _pair("resnet18", 3, 4)
        prod = prod.double().train()
        prod.compute_dtype = torch.float64
        return FusedTrainer(prod, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.25] * 3,
                            acc_dtype=torch.float64, clip_grad_norm=1.0, **kw)
    plain, scaled = trainer(), trainer(loss_scale=1024.0)
    l0, l1 = plain.step(raw, t), scaled.step(raw, t)
    assert float(l0) == float(l1)  # the reported loss is never scaled
    assert torch.allclose(plain.flat, scaled.flat, atol=1e-12, rtol=1e-9)
    scaled.forward_backward(raw, t)
    plain.forward_backward(raw, t)
    assert torch.allclose(scaled.gflat, 1024.0 * plain.gflat, atol=1e-9, rtol=1e-9)  # gradients carry S until the step
    dyn = trainer(loss_scale="dynamic", growth_interval=2)
    assert float(dyn.loss_scale) == 65536.0
    dyn.forward_backward(raw, t)
    dyn.gflat[7] = float("inf")
    before = dyn.flat.clone()
    dyn.optimizer_step()
    assert torch.equal(dyn.flat, before) and float(dyn.loss_scale) == 32768.0 and dyn.skipped_steps == 1
    dyn.step(raw, t)
    assert not torch.equal(dyn.flat, before) and float(dyn.loss_scale) == 32768.0
    dyn.step(raw, t)
    assert float(dyn.loss_scale) == 65536.0  # two clean steps: the scale grows back
    with pytest.raises(NotImplementedError):
        trainer(loss_scale="dynamic", cuda_graph=True)
    with pytest.raises(ValueError):
        trainer(loss_scale="auto")


@pytest.mark.parametrize("name,cin,hw,folded", [("mit_b0", 3, 64, 1), ("mit_b1", 4, 128, 1), ("mit_b0", 3, 64, 0)])
def test_segformer_backward_equals_oracle_autograd(monkeypatch, f64, name, cin, hw, folded):
    """SegFormer graph wiring (attention GEMM operand slicing, fp32 residual stream, LayerNorm chain,
    virtual-concat decoder, bilinear heads) against the reference-pinned functional oracle, in float64.
    folded = 1 (the default): the decoder applies linear_fuse's 1x1 conv in front of the resizes, composed with the four
    projections (SegFormer._decoder_folded_fwd) — the same function and the same parameter gradients as the reference's
    op order (folded = 0), to float64 rounding."""
    from gdl_b200 import ops as _ops
    from gdl_b200.engine import Act, Engine
    from gdl_b200.models.segformer import SegFormer
    from oracle import segformer as osf
    emu.install(monkeypatch)
    monkeypatch.setitem(_ops._HOST_OPTS, "decoder_folded", folded)
    k = 5
    torch.manual_seed(0)
    prod = SegFormer(name, in_channels=cin, num_classes=k, compute_dtype=torch.float64).double().train()
    with torch.no_grad():  # make every affine parameter / bias non-trivial
        for n_, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {n_: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n_ else v.clone())
          for n_, v in prod.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, hw, hw, generator=g).double()
    t = torch.randint(0, k, (2, hw, hw), generator=g)
    ref = osf.segformer_forward(sd, x, name, training=True)
    F.cross_entropy(ref, t).backward()

    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    with torch.no_grad():
        xin = emu.normalize_to_nhwc(x, True, torch.float64, 8)
        logits = prod.run(eng, Act(xin, needs_grad=False))
    assert (eng.saved_segformer.projs is None) == bool(folded)  # the route under test really ran
    assert torch.allclose(logits.permute(0, 3, 1, 2), ref, atol=1e-9, rtol=1e-9)
    d = logits.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.cross_entropy(d, t).backward()
    with torch.no_grad():
        prod.backward(eng, d.grad.permute(0, 2, 3, 1).contiguous())
    for n_, p in prod.named_parameters():
        got, want = eng.param_grads[id(p)], sd[n_].grad
        err = (got - want).abs().max() / (want.abs().max() + 1e-30)
        # parameters whose true gradient is ~0 (biases in front of a train-mode BatchNorm) only carry noise
        assert err < 1e-7 or want.abs().max() < 1e-12, f"{n_}: {err}"
    assert torch.allclose(prod.decoder.linear_fuse[1].running_mean, sd["decoder.linear_fuse.1.running_mean"], atol=1e-12)


def test_segformer_stochastic_layers_with_supplied_draws(monkeypatch, f64):
    """DropPath (per-sample factors on both branches of every block) and the decoder's Dropout2d, with the random
    draws supplied to product and oracle alike: forward and every parameter gradient agree in float64; in eval mode the
    layers are identities; with rates set and no supplied draws the training forward is stochastic."""
    from gdl_b200.engine import Act, Engine
    from gdl_b200.models.segformer import MIT_CFG, SegFormer
    from oracle import segformer as osf
    emu.install(monkeypatch)
    name, cin, hw, k, b = "mit_b0", 3, 64, 4, 3
    torch.manual_seed(0)
    prod = SegFormer(name, in_channels=cin, num_classes=k, compute_dtype=torch.float64, drop_path_rate=0.1,
                     dropout_ratio=0.1).double().train()
    nblk = sum(MIT_CFG[name][2])
    assert len(prod.drop_path_rates) == nblk and prod.drop_path_rates[0] == 0.0 and abs(prod.drop_path_rates[-1] - 0.1) < 1e-7
    g = torch.Generator().manual_seed(1)
    masks = []
    for i in range(nblk):
        keep = 1.0 - 0.1 * i / (nblk - 1)
        masks.append(tuple((torch.rand(b, generator=g) < keep).double() / keep for _ in range(2)))
    masks[3] = (torch.tensor([0.0, 1.0, 2.0]).double(), torch.tensor([1.5, 0.0, 0.0]).double())
    emb = MIT_CFG[name][3]
    dmask = (torch.rand(b, emb, generator=g) < 0.8).double() / 0.8
    prod.drop_path_masks, prod.dropout_mask = masks, dmask
    sd = {n_: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n_ else v.clone())
          for n_, v in prod.state_dict().items()}
    x = torch.randn(b, cin, hw, hw, generator=g).double()
    t = torch.randint(0, k, (b, hw, hw), generator=g)
    ref = osf.segformer_forward(sd, x, name, training=True, drop_path=masks, dropout_mask=dmask)
    F.cross_entropy(ref, t).backward()
    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    with torch.no_grad():
        xin = emu.normalize_to_nhwc(x, True, torch.float64, 8)
        logits = prod.run(eng, Act(xin, needs_grad=False))
    assert torch.allclose(logits.permute(0, 3, 1, 2), ref, atol=1e-9, rtol=1e-9)
    d = logits.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.cross_entropy(d, t).backward()
    with torch.no_grad():
        prod.backward(eng, d.grad.permute(0, 2, 3, 1).contiguous())
    for n_, p in prod.named_parameters():
        got, want = eng.param_grads[id(p)], sd[n_].grad
        err = (got - want).abs().max() / (want.abs().max() + 1e-30)
        assert err < 1e-7 or want.abs().max() < 1e-12, f"{n_}: {err}"
    # eval: identities, whatever the rates
    prod.drop_path_masks = prod.dropout_mask = None
    with torch.no_grad():
        e1 = prod.run(Engine(torch.float64, training=False, acc_dtype=torch.float64), Act(xin, needs_grad=False))
        plain = osf.segformer_forward({n_: v.detach() for n_, v in sd.items()}, x, name, training=False)
    assert torch.allclose(e1.permute(0, 3, 1, 2), plain, atol=1e-9, rtol=1e-9)
    # training with rates set and no supplied draws: two passes differ (the layers really draw)
    with torch.no_grad():
        torch.manual_seed(1)
        a = prod.run(Engine(torch.float64, training=True, acc_dtype=torch.float64), Act(xin, needs_grad=False))
        torch.manual_seed(2)
        c = prod.run(Engine(torch.float64, training=True, acc_dtype=torch.float64), Act(xin, needs_grad=False))
    assert not torch.allclose(a, c)


@pytest.mark.parametrize("name,cin", [("mit_b0", 3), ("mit_b1", 6)])
def test_dynamic_mix_transformer_backward_equals_oracle_autograd(monkeypatch, f64, name, cin):
    """`use_dynamic_encoder=True` (DynamicMixTransformer / DynamicChannelEmbed, mix_transformer.py:762-934): the band-wise
    layers as block-diagonal dense convolutions + the channel-pool kernels + torch autograd for the tiny weight
    construction, against the oracle (pinned to the reference's own module) — forward and every parameter gradient."""
    from gdl_b200.engine import Act, Engine
    from gdl_b200.models.segformer import SegFormer
    from oracle import segformer as osf
    emu.install(monkeypatch)
    k, hw = 4, 64
    torch.manual_seed(0)
    prod = SegFormer(name, in_channels=3, num_classes=k, use_dynamic_encoder=True, compute_dtype=torch.float64).double().train()
    assert not hasattr(prod.encoder, "patch_embed1") and "encoder.dynamic_patch_embed1.spatial_conv.weight" in prod.state_dict()
    with torch.no_grad():
        for n_, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {n_: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n_ else v.clone())
          for n_, v in prod.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, hw, hw, generator=g).double()
    t = torch.randint(0, k, (2, hw, hw), generator=g)
    ref = osf.segformer_forward(sd, x, name, training=True)
    F.cross_entropy(ref, t).backward()
    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    with torch.no_grad():
        xin = emu.normalize_to_nhwc(x, True, torch.float64, 8)
        logits = prod.run(eng, Act(xin, needs_grad=False), cin)
    assert torch.allclose(logits.permute(0, 3, 1, 2), ref, atol=1e-9, rtol=1e-9)
    d = logits.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.cross_entropy(d, t).backward()
    with torch.no_grad():
        prod.backward(eng, d.grad.permute(0, 2, 3, 1).contiguous())
    for n_, p in prod.named_parameters():
        got, want = eng.param_grads[id(p)], sd[n_].grad
        err = (got - want).abs().max() / (want.abs().max() + 1e-30)
        assert err < 1e-7 or want.abs().max() < 1e-12, f"{n_}: {err}"
    # the same model accepts another band count without any change (what "dynamic" means) through the public forward
    prod.eval()
    x5 = torch.randn(1, 5, hw, hw, generator=g).double()
    with torch.no_grad():
        want5 = osf.segformer_forward({n_: v.detach() for n_, v in sd.items()}, x5, name)
        assert torch.allclose(prod(x5).double(), want5, atol=1e-5, rtol=1e-5)  # the public forward takes the image as fp32


def test_upernet_aux_head_dropout_with_supplied_draw(monkeypatch, f64):
    """FCNHead's Dropout2d (fcn_head.py:69-83) on the engine tape: y = x * m[n][c] forward, the same scaling backward."""
    from gdl_b200.engine import Act, Engine
    emu.install(monkeypatch)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 6, 16, generator=g).double()
    m = (torch.rand(3, 16, generator=g) < 0.7).double() / 0.7
    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    a = Act(x)
    y = eng.dropout2d(a, 0.3, m)
    assert torch.equal(y.t, x * m.view(3, 1, 1, 16))
    gy = torch.randn(3, 5, 6, 16, generator=g).double()
    y.gsrcs.append((gy, 0))
    eng.backward()
    assert torch.equal(eng.collect_grad(a), gy * m.view(3, 1, 1, 16))
    ev = Engine(torch.float64, training=False, acc_dtype=torch.float64)
    assert ev.dropout2d(a, 0.3) is a and eng.dropout2d(a, 0.0) is a
    drawn = eng.dropout2d(Act(torch.ones(4, 2, 2, 64).double()), 0.5).t
    vals = set(drawn.unique().tolist())
    assert vals <= {0.0, 2.0} and len(vals) == 2 and torch.equal(drawn[:, 0, 0], drawn[:, 1, 1])  # whole planes


def test_upernet_backward_equals_oracle_autograd(monkeypatch, f64):
    """MultiLevelNeck + UperNet + heads wiring (PPM pooling, top-down adds, virtual concats, two logit maps)."""
    from gdl_b200.engine import Act, Engine
    from gdl_b200.models.upernet import UperNetSegmentor
    from oracle import upernet as ou
    emu.install(monkeypatch)
    e, ch, k = 96, 64, 5
    torch.manual_seed(0)
    prod = UperNetSegmentor(e, ch, k, compute_dtype=torch.float64).double().train()
    with torch.no_grad():
        for _, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {n_: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n_ else v.clone())
          for n_, v in prod.state_dict().items()}
    assert set(sd) == set(ou.init_state_dict(e, ch, k))
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(2, e, 12, 12, generator=g).double().requires_grad_(True) for _ in range(4)]
    t = torch.randint(0, k, (2, 168, 168), generator=g)
    ro, ra = ou.upernet_forward(sd, feats, (168, 168), training=True)
    (F.cross_entropy(ro, t) + 0.4 * F.cross_entropy(ra, t)).backward()

    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    with torch.no_grad():
        acts = [Act(f.detach().permute(0, 2, 3, 1).contiguous(), needs_grad=True) for f in feats]
        o, a = prod.run(eng, acts, (168, 168))
    assert torch.allclose(o.permute(0, 3, 1, 2), ro, atol=1e-9) and torch.allclose(a.permute(0, 3, 1, 2), ra, atol=1e-9)
    do = o.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    da = a.detach().permute(0, 3, 1, 2).clone().requires_grad_(True)
    (F.cross_entropy(do, t) + 0.4 * F.cross_entropy(da, t)).backward()
    with torch.no_grad():
        prod.backward(eng, do.grad.permute(0, 2, 3, 1).contiguous(), da.grad.permute(0, 2, 3, 1).contiguous())
        fg = [eng.collect_grad(x) for x in acts]
    for n_, p in prod.named_parameters():
        got, want = eng.param_grads[id(p)], sd[n_].grad
        err = (got - want).abs().max() / (want.abs().max() + 1e-30)
        assert err < 1e-7 or want.abs().max() < 1e-12, f"{n_}: {err}"
    for f, gf in zip(feats, fg):  # gradients reach the encoder maps too (un-frozen encoder case)
        assert torch.allclose(gf.permute(0, 3, 1, 2), f.grad, atol=1e-10)


def test_dofa_encoder_forward_equals_oracle(monkeypatch, f64):
    """DOFA-v2 encoder host logic (wave embedding, weight-generator layer, dynamic patch conv through im2col, token
    assembly, ViT blocks with LayerScale, taps) against the oracle restatement, in float64."""
    from gdl_b200.models.dofa import DOFAv2
    from oracle import dofa as od
    emu.install(monkeypatch)
    e, depth, heads, img = 96, 3, 3, 70
    sd = {k: v.double() for k, v in od.init_state_dict(e, depth, img, seed=3, ls_init=0.5).items()}
    prod = DOFAv2("dofa_base", img, 14, e, depth, heads, out_indices=[0, 2], compute_dtype=torch.float64).double()
    assert set(prod.state_dict()) == set(sd)
    prod.load_state_dict(sd)
    for p in prod.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 5, img, img, generator=g).double()
    wl = torch.tensor([0.49, 0.56, 0.665, 0.842, 1.61]).double()
    want = od.dofa_forward(sd, x, wl, e, depth, heads, out_indices=(0, 2))
    got = prod(x, wl)
    assert len(got) == len(want) == 2
    for a, b in zip(got, want):
        assert a.shape == b.shape
        err = (a - b).abs().max() / b.abs().max()
        assert err < 1e-5, err  # the fp32 token stream / fp32 generator intermediates bound this, not the wiring
    with pytest.raises(ValueError):
        prod(x, torch.stack([wl, wl * 1.1]))


def test_dofa_convert_patch_to_16(monkeypatch, f64):
    """DOFAv2(convert_patch_to_16=True): bicubic 14 -> 16 resampling of the generated kernels, stride-16 embedding; frozen
    route (kernels) and trainable route (autograd through the resampling) against the oracle."""
    from gdl_b200.engine import Engine
    from gdl_b200.models.dofa import DOFAv2
    from oracle import dofa as od
    emu.install(monkeypatch)
    torch.manual_seed(0)
    enc = DOFAv2("dofa_base", 64, 14, 32, 2, 2, out_indices=[0, 1], convert_patch_to_16=True, drop_path_rate=0.0,
                 compute_dtype=torch.float64).double()
    assert enc.num_patches == 16 and enc.pos_embed.shape == (1, 17, 32)
    with torch.no_grad():
        for n_, p in enc.named_parameters():
            if "ls1" in n_ or "ls2" in n_:
                p.fill_(0.5)
    sd = {n_: v.detach().clone().requires_grad_(v.is_floating_point() and n_ != "pos_embed") for n_, v in enc.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 64, 64, generator=g).double()
    wl = torch.tensor([0.665, 0.56, 0.49]).double()
    want = od.dofa_forward(sd, x, wl, 32, 2, 2, out_indices=(0, 1), convert_to_16=True)
    with torch.no_grad():
        got = enc(x, wl)
    for a, b in zip(got, want):
        # (the frozen route evaluates the wavelength sin/cos in fp32 like the reference; the float64 oracle does not)
        assert a.shape == b.shape == (2, 32, 4, 4) and (a - b).abs().max() < 1e-5 * b.abs().max()
    # trainable route: gradients through the resampled kernels reach the weight generator
    (want[0].sum() + 2 * want[1].sum()).backward()
    eng = Engine(torch.float64, training=True, acc_dtype=torch.float64)
    with torch.no_grad():
        feats = enc.run_train(eng, emu.normalize_to_nhwc(x, True, torch.float64, 8), 3, wl)
        for f, wgt in zip(feats, (1.0, 2.0)):
            f.gsrcs.append((torch.full_like(f.t, wgt), 0))
        enc.backward(eng)
    for n_ in ("patch_embed.weight_generator.fc_weight.weight", "patch_embed.fclayer.w1.weight", "blocks.0.attn.qkv.weight"):
        p = dict(enc.named_parameters())[n_]
        err = (eng.param_grads[id(p)] - sd[n_].grad).abs().max() / sd[n_].grad.abs().max()
        assert err < 1e-6, (n_, err)


def test_dofa_unfrozen_encoder_is_rejected(monkeypatch):
    from gdl_b200.models.dofa import DOFAv2
    emu.install(monkeypatch)
    enc = DOFAv2("dofa_base", 28, 14, 32, 1, 2)
    with pytest.raises(NotImplementedError):
        enc.forward_features(torch.zeros(1, 3, 28, 28), torch.tensor([0.6, 0.5, 0.4]))


def test_dofa_segmentation_model_train_step_equals_oracle(monkeypatch, f64):
    """DOFASegmentationModel (frozen encoder -> neck -> UperNet -> heads) through its public forward + autograd."""
    from gdl_b200.models.dofa import DOFASegmentationModel
    from oracle import dofa as od, upernet as ou
    emu.install(monkeypatch)
    torch.manual_seed(0)
    m = DOFASegmentationModel("dofa_base", (56, 56), ["encoder"], 4, compute_dtype=torch.float64).double().train()
    m.acc_dtype = torch.float64
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "ls1" in n_ or "ls2" in n_:
                p.fill_(0.3)
    assert not any(p.requires_grad for p in m.encoder.parameters())
    sd = {n_: (v.detach().clone().requires_grad_(True)
               if v.is_floating_point() and "running" not in n_ and not n_.startswith("encoder.") else v.clone())
          for n_, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 56, 56, generator=g).double()
    wl = torch.tensor([0.665, 0.56, 0.49]).double()
    t = torch.randint(0, 4, (2, 56, 56), generator=g)
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    with torch.no_grad():
        feats = od.dofa_forward(enc_sd, x, wl)
        mine = m.encoder(x, wl)
    for a, b in zip(mine, feats):
        assert (a - b).abs().max() < 1e-5 * b.abs().max()
    # the decoder's BatchNorms see 2..32 samples per channel here and amplify the encoder's fp32-stream rounding by
    # orders of magnitude, so the head is compared on the SAME encoder maps
    feats = [f.clone() for f in mine]
    ro, ra = ou.upernet_forward({k: v for k, v in sd.items() if not k.startswith("encoder.")}, feats, (56, 56), training=True)
    (F.cross_entropy(ro, t) + 0.4 * F.cross_entropy(ra, t)).backward()
    out = m(x, wl)
    assert (out.out - ro).abs().max() < 1e-8 * ro.abs().max() and (out.aux - ra).abs().max() < 1e-8 * ra.abs().max()
    (F.cross_entropy(out.out, t) + 0.4 * F.cross_entropy(out.aux, t)).backward()
    for n_, p in m.named_parameters():
        if n_.startswith("encoder."):
            assert p.grad is None
            continue
        want = sd[n_].grad
        err = (p.grad - want).abs().max() / (want.abs().max() + 1e-30)
        assert err < 1e-6 or want.abs().max() < 1e-12, f"{n_}: {err}"


@pytest.mark.parametrize("drop_path", [False, True])
def test_dofa_trainable_encoder_backward_equals_oracle_autograd(monkeypatch, f64, drop_path):
    """Un-frozen DOFA encoder: hand-written backward of the ViT blocks (LayerScale / DropPath factors, GELU, attention,
    LayerNorm), the token glue, the dynamic patch embedding and — through torch autograd — the weight generator, against
    the oracle's autograd for every parameter of the model."""
    from gdl_b200.models.dofa import DOFASegmentationModel
    from oracle import dofa as od, upernet as ou
    emu.install(monkeypatch)
    torch.manual_seed(0)
    m = DOFASegmentationModel("dofa_base", (56, 56), None, 4, compute_dtype=torch.float64).double().train()
    m.acc_dtype = torch.float64
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "ls1" in n_ or "ls2" in n_:
                p.copy_(0.3 + 0.1 * torch.rand_like(p))
    assert m.encoder.trainable()
    g = torch.Generator().manual_seed(1)
    b = 2
    if drop_path:  # timm DropPath with the draw supplied: some samples dropped, the others scaled by 1 / keep
        masks = []
        for i in range(12):
            keep = 1.0 - 0.05 * i
            masks.append(tuple((torch.rand(b, generator=g) < keep).double() / keep for _ in range(2)))
        masks[11] = (torch.tensor([0.0, 1.25]).double(), torch.tensor([1.25, 0.0]).double())
        m.encoder.drop_path_masks = masks
    else:
        masks = None
        m.encoder.drop_path_rates = [0.0] * 12
    sd = {n_: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n_ and n_ != "encoder.pos_embed"
               else v.clone()) for n_, v in m.state_dict().items()}
    x = torch.randn(b, 3, 56, 56, generator=g).double()
    wl = torch.tensor([0.665, 0.56, 0.49]).double()
    t = torch.randint(0, 4, (b, 56, 56), generator=g)
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    feats = od.dofa_forward(enc_sd, x, wl, drop_path=masks)
    ro, ra = ou.upernet_forward({k: v for k, v in sd.items() if not k.startswith("encoder.")}, feats, (56, 56), training=True)
    (F.cross_entropy(ro, t) + 0.4 * F.cross_entropy(ra, t)).backward()
    out = m(x, wl)
    assert (out.out - ro).abs().max() < 1e-7 * ro.abs().max() and (out.aux - ra).abs().max() < 1e-7 * ra.abs().max()
    (F.cross_entropy(out.out, t) + 0.4 * F.cross_entropy(out.aux, t)).backward()
    checked = 0
    for n_, p in m.named_parameters():
        if not p.requires_grad:
            continue
        want = sd[n_].grad
        if want is None:  # encoder.norm is never applied (reference quirk) and has no gradient on either side
            assert p.grad is None, n_
            continue
        assert p.grad is not None, n_
        err = (p.grad - want).abs().max() / (want.abs().max() + 1e-30)
        assert err < 1e-5 or want.abs().max() < 1e-12, f"{n_}: {err}"
        checked += 1
    assert checked > 200


def test_dofa_fused_trainer_matches_autograd_route(monkeypatch, f64):
    """FusedTrainer's `fused_train` hook (frozen encoder, two logit maps, 0.4 aux weight) leaves the same gradients
    in the flat buffer as the autograd route checked above, and only the trainable half is in the Adam buffers."""
    from gdl_b200.models.dofa import DOFASegmentationModel
    from gdl_b200.trainer import FusedTrainer
    emu.install(monkeypatch)
    torch.manual_seed(0)
    m = DOFASegmentationModel("dofa_base", (56, 56), ["encoder"], 4, compute_dtype=torch.float64).double().train()
    m.acc_dtype = torch.float64
    m.wavelengths = torch.tensor([0.665, 0.56, 0.49]).double()
    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 256, (2, 56, 56, 3), generator=g, dtype=torch.uint8)
    t = torch.randint(0, 4, (2, 56, 56), generator=g)
    mean, std = [0.5] * 3, [0.25] * 3
    x = ((raw.double() / 255.0).permute(0, 3, 1, 2) - 0.5) / 0.25
    out = m(x, m.wavelengths)
    (F.cross_entropy(out.out, t) + 0.4 * F.cross_entropy(out.aux, t)).backward()
    want = {n_: p.grad.clone() for n_, p in m.named_parameters() if p.requires_grad}
    for mod in m.modules():  # the autograd route above already advanced the running statistics once
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.reset_running_stats()
    tr = FusedTrainer(m, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=mean, std=std, acc_dtype=torch.float64)
    assert tr.flat.numel() == sum((p.numel() + 3) // 4 * 4 for n_, p in m.named_parameters() if not n_.startswith("encoder."))
    loss = tr.forward_backward(raw, t)
    # (the autograd route rounds the image to fp32 before the normalise kernel: ~1e-8 relative input difference)
    assert abs(float(loss) - float(F.cross_entropy(out.out, t) + 0.4 * F.cross_entropy(out.aux, t))) < 1e-7
    for n_, p in m.named_parameters():
        if p.requires_grad:
            err = (p.grad - want[n_]).abs().max() / (want[n_].abs().max() + 1e-30)
            assert err < 1e-4 or want[n_].abs().max() < 1e-12, f"{n_}: {err}"


def test_dofa_fused_trainer_with_trainable_encoder(monkeypatch, f64):
    """FusedTrainer on an un-frozen DOFA model: every parameter (encoder included) lives in the flat buffers and gets
    the gradient of the autograd route; one Adam step moves encoder weights."""
    from gdl_b200.models.dofa import DOFASegmentationModel
    from gdl_b200.trainer import FusedTrainer
    emu.install(monkeypatch)
    torch.manual_seed(0)
    m = DOFASegmentationModel("dofa_base", (56, 56), None, 4, compute_dtype=torch.float64).double().train()
    m.acc_dtype = torch.float64
    m.encoder.drop_path_rates = [0.0] * 12
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "ls1" in n_ or "ls2" in n_:
                p.fill_(0.3)
    m.wavelengths = torch.tensor([0.665, 0.56, 0.49]).double()
    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 256, (2, 56, 56, 3), generator=g, dtype=torch.uint8)
    t = torch.randint(0, 4, (2, 56, 56), generator=g)
    x = ((raw.double() / 255.0).permute(0, 3, 1, 2) - 0.5) / 0.25
    out = m(x, m.wavelengths)
    (F.cross_entropy(out.out, t) + 0.4 * F.cross_entropy(out.aux, t)).backward()
    want = {n_: p.grad.clone() for n_, p in m.named_parameters() if p.grad is not None}
    assert any(n_.startswith("encoder.blocks.") for n_ in want) and any("weight_generator" in n_ for n_ in want)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.reset_running_stats()
    tr = FusedTrainer(m, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.25] * 3,
                      acc_dtype=torch.float64)
    assert tr.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in m.parameters() if p.requires_grad)
    before = m.encoder.blocks[3].attn.qkv.weight.detach().clone()
    tr.forward_backward(raw, t)
    for n_, p in m.named_parameters():
        if n_ in want:
            err = (p.grad - want[n_]).abs().max() / (want[n_].abs().max() + 1e-30)
            assert err < 1e-4 or want[n_].abs().max() < 1e-12, f"{n_}: {err}"
    tr.optimizer_step()
    assert not torch.equal(m.encoder.blocks[3].attn.qkv.weight, before)


def test_sliding_window_inference_equals_window_sum_of_oracle(monkeypatch, f64):
    """SlidingWindowSegmenter host logic (window grid incl. the border-flush windows, zero padding of small rasters,
    overlap blending by logit sum, argmax) against a direct restatement on the oracle model."""
    from gdl_b200.inference import SlidingWindowSegmenter, window_origins
    emu.install(monkeypatch)
    assert window_origins(100, 64, 32) == [0, 32, 36] and window_origins(64, 64, 32) == [0] and window_origins(40, 64, 32) == [0]
    ora, prod = _pair("resnet18", 3, 4)
    ora, prod = ora.double().eval(), prod.double().eval()
    prod.compute_dtype = torch.float64
    mean, std = [0.4, 0.5, 0.6], [0.2, 0.25, 0.3]
    g = torch.Generator().manual_seed(9)
    for (h, w) in ((100, 150), (40, 64)):
        raster = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
        seg = SlidingWindowSegmenter(prod, tile=64, stride=32, batch=3, mean=mean, std=std)
        got = seg.logits(raster)
        hp, wp = max(h, 64), max(w, 64)
        padded = torch.zeros(hp, wp, 3, dtype=torch.uint8)
        padded[:h, :w] = raster
        x = ((padded.double() / 255.0) - torch.tensor(mean).double()) / torch.tensor(std).double()
        want = torch.zeros(hp, wp, 4, dtype=torch.float64)
        n = 0
        with torch.no_grad():
            for y in window_origins(hp, 64, 32):
                for xx in window_origins(wp, 64, 32):
                    lo = ora(x[y:y + 64, xx:xx + 64].permute(2, 0, 1).unsqueeze(0))[0]
                    want[y:y + 64, xx:xx + 64] += lo.permute(1, 2, 0)
                    n += 1
        assert seg.windows_done == n
        assert torch.allclose(got.double(), want[:h, :w], atol=1e-5, rtol=1e-5)  # the logit accumulator is fp32
        cls = seg.predict(raster)
        assert cls.dtype == torch.uint8 and cls.shape == (h, w)
        margin = want[:h, :w].topk(2, dim=2).values
        sure = (margin[..., 0] - margin[..., 1]) > 1e-4
        assert torch.equal(cls[sure].long(), want[:h, :w].argmax(2)[sure])


def test_trainer_gradient_buckets_cover_the_flat_buffer_once(monkeypatch):
    """The overlapped all-reduce (FusedTrainer.overlap_allreduce) reduces contiguous buckets of the flat gradient buffer in
    backward order: together they must cover every element exactly once, each parameter must belong to exactly one bucket,
    and the first bucket must hold the layers whose backward runs first (the head / decoder = the tail of the buffer)."""
    from gdl_b200.trainer import FusedTrainer, gradient_buckets
    emu.install(monkeypatch)
    _, prod = _pair("resnet18", 3, 4)
    tr = FusedTrainer(prod, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-2, mean=[0.5] * 3, std=[0.25] * 3)
    assert tr._buckets == [] and not tr.overlap_allreduce  # single process: nothing to reduce
    total = tr.flat.numel()
    for nb in (1, 2, 3, 7):
        buckets = gradient_buckets([p.numel() for p in tr.params], [id(p) for p in tr.params], nb)
        assert 1 <= len(buckets) <= nb
        assert [b[0] for b in buckets] == sorted((b[0] for b in buckets), reverse=True)       # backward order
        assert buckets[-1][0] == 0 and buckets[0][1] == total                                 # covers the whole buffer ...
        assert all(buckets[i][0] == buckets[i + 1][1] for i in range(len(buckets) - 1))       # ... contiguously, no overlap
        all_ids = [i for b in buckets for i in b[2]]
        assert sorted(all_ids) == sorted(id(p) for p in tr.params)                            # every parameter exactly once
        assert id(prod.segmentation_head[0].weight) in buckets[0][2]                          # head: first to finish its backward
        assert id(prod.encoder.conv1.weight) in buckets[-1][2]                                # stem: last
        # a parameter's slot lies inside its bucket
        off = 0
        for p in tr.params:
            b = next(b for b in buckets if id(p) in b[2])
            assert b[0] <= off and off + p.numel() <= b[1]
            off += (p.numel() + 3) // 4 * 4
    # one parameter larger than a whole share: it closes its bucket, the walk skips the bounds it ran over
    bs = gradient_buckets([4, 100, 4, 4], list("abcd"), 3)
    assert [(a, b, sorted(k)) for a, b, k in bs] == [(104, 112, ["c", "d"]), (0, 104, ["a", "b"])]
    assert gradient_buckets([], [], 3) == []
