"""GPU: SegFormer (MixTransformer + MLP decoder) parity against the reference-pinned oracle
(oracle/segformer.py == the reference's own modules, see tests/test_oracle_cpu.py).

Same bar as the UNet++ model tests: the 16-bit product may deviate from the fp32 oracle by at most
2.5x (logits) / 3x (gradients) what the oracle itself deviates when run under torch.autocast."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def _setup(name, cin, k, dtype=torch.bfloat16, seed=0):
    from gdl_b200.models.segformer import SegFormer
    torch.manual_seed(seed)
    prod = SegFormer(name, in_channels=cin, num_classes=k, compute_dtype=dtype).cuda()
    with torch.no_grad():
        for _, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return prod


def _oracle_sd(prod):
    return {n: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in n else v.clone())
            for n, v in prod.state_dict().items()}


@pytest.fixture
def decoder_route():
    """set the "decoder_folded" host option for one test, restore it afterwards"""
    from gdl_b200 import ops
    old = ops.option("decoder_folded")
    yield lambda v: ops.set_option("decoder_folded", v)
    ops.set_option("decoder_folded", old)


@pytest.mark.parametrize("name,cin,k,hw,dtype,folded", [
    ("mit_b0", 3, 5, 128, torch.bfloat16, 1),
    ("mit_b0", 3, 5, 64, torch.bfloat16, 1),      # 4 keys per image: exercises the key padding to 16
    ("mit_b2", 4, 5, 128, torch.bfloat16, 1),
    ("mit_b1", 6, 2, 128, torch.float16, 1),
    ("mit_b0", 3, 5, 128, torch.bfloat16, 0),     # the reference's op order in the decoder (resize, concat, 4*emb -> emb conv)
    ("mit_b2", 4, 5, 128, torch.bfloat16, 0),
])
def test_train_step_parity(cuda, decoder_route, name, cin, k, hw, dtype, folded):
    """folded = 1 (default): linear_fuse's 1x1 conv runs in front of the resizes, composed with linear_c1..4
    (SegFormer._decoder_folded_fwd); both routes meet the same bar against the fp32 oracle."""
    decoder_route(folded)
    train_step_parity(name, cin, k, hw, dtype)


def train_step_parity(name, cin, k, hw, dtype):
    """one training step (forward, CE, backward) through the nn.Module surface against the fp32 oracle; the bar is what the
    reference stack itself deviates under torch.autocast on the same tiles (also used at the BASELINE shapes:
    tests/test_baseline_shapes_gpu.py)"""
    from oracle import segformer as osf
    prod = _setup(name, cin, k, dtype)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, cin, hw, hw, generator=g).cuda()
    t = torch.randint(0, k, (4, hw, hw), generator=g).cuda()
    sd = _oracle_sd(prod)
    ref = osf.segformer_forward(sd, x, name, training=True)
    ref_loss = F.cross_entropy(ref, t)
    ref_loss.backward()
    sd_ac = _oracle_sd(prod)
    with torch.autocast("cuda", dtype=dtype):
        ac = osf.segformer_forward(sd_ac, x, name, training=True)
    ac_loss = F.cross_entropy(ac.float(), t)
    ac_loss.backward()

    prod.train()
    logits = prod(x)
    assert logits.shape == ref.shape and logits.dtype == torch.float32
    loss = F.cross_entropy(logits, t)
    loss.backward()
    e_prod, e_ac = _rel(logits, ref), _rel(ac, ref)
    print(f"[{name} {dtype} {hw}] logits rel err: product {e_prod:.4f}, autocast reference {e_ac:.4f}")
    assert e_prod < max(2.5 * e_ac, 5e-3)
    assert abs(loss.item() - ref_loss.item()) < max(2.5 * abs(ac_loss.item() - ref_loss.item()), 5e-3)
    rows = []
    for n, p in prod.named_parameters():
        want = sd[n].grad
        if want.abs().max() < 1e-9:  # analytically zero (bias in front of train-mode BN): only noise
            continue
        rows.append((n, _rel(p.grad, want), _rel(sd_ac[n].grad, want)))
    worst = max(r[1] / max(r[2], 2e-3) for r in rows)
    print(f"[{name} {dtype} {hw}] worst grad err ratio vs autocast: {worst:.2f}")
    for n, ep, ea in rows[:4] + rows[-4:]:
        print(f"    {n:45s} product {ep:.4f} autocast {ea:.4f}")
    for n, ep, ea in rows:
        assert ep < max(3.0 * ea, 2e-2), f"{n}: product {ep:.4f} vs autocast {ea:.4f}"


def test_eval_forward_and_features(cuda):
    from oracle import segformer as osf
    prod = _setup("mit_b2", 3, 5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 256, 256, generator=g).cuda()
    sd = {k: v.detach() for k, v in prod.state_dict().items()}
    prod.eval()
    with torch.no_grad():
        ref = osf.segformer_forward(sd, x, "mit_b2", training=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = osf.segformer_forward(sd, x, "mit_b2", training=False).float()
        out = prod(x)
    assert _rel(out, ref) < max(2.5 * _rel(ac, ref), 5e-3)
    agree, agree_ac = (out.argmax(1) == ref.argmax(1)).float().mean().item(), (ac.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"argmax agreement with fp32 oracle: product {agree:.4f}, autocast reference {agree_ac:.4f}")
    assert agree >= agree_ac - 0.005


def test_fused_trainer_reduces_loss(cuda):
    from gdl_b200.ops import LossSpec
    from gdl_b200.trainer import FusedTrainer
    prod = _setup("mit_b0", 3, 5, seed=3).train()
    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (4, 4, 4), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    raw = (t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 128, 128, 3), generator=g).cuda()).to(torch.uint8)
    tr = FusedTrainer(prod, LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.2] * 3,
                      clip_grad_norm=1.0)
    losses = [tr.step(raw, t).item() for _ in range(15)]
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < 0.7 * losses[0]


@pytest.mark.parametrize("fused", [0, 1])
def test_cuda_graph_step_equals_eager_step_bitwise(cuda, fused):
    """SegFormer's fused step (attention three-kernel path or the single fused kernel, LayerNorm / DWConv parameter
    gradients, batched dK / dV weight gradients, gradient clipping): the CUDA-graph replay, a second eager run and the
    first agree bit for bit — every cross-block sum is ordered (gdl_set_workspace)."""
    from gdl_b200 import ops
    from gdl_b200.ops import LossSpec
    from gdl_b200.trainer import FusedTrainer
    assert ops.deterministic()
    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (4, 4, 4), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    raw = (t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 128, 128, 3), generator=g).cuda()).to(torch.uint8)
    old = ops.option("sra_fused")
    ops.set_option("sra_fused", fused)
    try:
        losses, flats = {}, {}
        for mode in ("eager", "eager2", "graph"):
            prod = _setup("mit_b1", 3, 5, seed=3).train()
            tr = FusedTrainer(prod, LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.2] * 3,
                              clip_grad_norm=1.0, cuda_graph=mode == "graph")
            losses[mode] = [tr.step(raw, t).item() for _ in range(5)]
            flats[mode] = tr.flat.clone()
    finally:
        ops.set_option("sra_fused", old)
    print("eager", losses["eager"], "graph", losses["graph"])
    assert losses["eager"] == losses["eager2"] and torch.equal(flats["eager"], flats["eager2"])
    assert losses["eager"] == losses["graph"] and torch.equal(flats["eager"], flats["graph"])
