"""GPU: DynamicMixTransformer (`use_dynamic_encoder=True`, mix_transformer.py:762-934): the channel-pool / ReLU-backward
kernels against torch, the reference's golden logits, and a train step under the tolerance rule of test_segformer_gpu.py.
(Sorts last on purpose: written after the round's GPU budget was spent; forward and every parameter gradient are pinned on
CPU in float64 by tests/test_engine_host_logic_cpu.py::test_dynamic_mix_transformer_backward_equals_oracle_autograd.)"""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from test_segformer_gpu import _oracle_sd, _rel

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("bands,e,dtype", [(3, 64, torch.bfloat16), (6, 32, torch.float16), (1, 64, torch.bfloat16), (16, 32, torch.bfloat16)])
def test_channel_pool_kernels(cuda, bands, e, dtype):
    from gdl_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(bands)
    n, h, w = 2, 19, 23
    xw = torch.randn(n, h, w, bands * e, generator=g, device="cuda").to(dtype)
    sc = torch.randn(n, h, w, 16, generator=g, device="cuda") * 2
    out, attn = ops.channel_pool_fwd(xw, sc, bands)
    a = torch.softmax(sc[..., :bands], -1)
    want = (xw.float().view(n, h, w, bands, e) * a.unsqueeze(-1)).sum(3)
    assert (attn[..., :bands] - a).abs().max() < 1e-6 and not attn[..., bands:].any()
    ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    assert (out.float() - want).abs().max() <= ulp * want.abs().max().clamp_min(1.0)
    d = torch.randn(n, h, w, e, generator=g, device="cuda").to(dtype)
    dxw, ds = ops.channel_pool_bwd(d, xw, attn, bands)
    want_dxw = (a.unsqueeze(-1) * d.float().unsqueeze(3)).reshape(n, h, w, bands * e)
    da = (xw.float().view(n, h, w, bands, e) * d.float().unsqueeze(3)).sum(-1)
    want_ds = a * (da - (a * da).sum(-1, keepdim=True))
    assert (dxw.float() - want_dxw).abs().max() <= ulp * want_dxw.abs().max().clamp_min(1.0)
    assert (ds[..., :bands].float() - want_ds).abs().max() <= ulp * want_ds.abs().max().clamp_min(1.0) + 1e-5
    assert not ds[..., bands:].any()
    y = torch.randn(n, h, w, 48, generator=g, device="cuda").to(dtype)
    dy = torch.randn(n, h, w, 48, generator=g, device="cuda").to(dtype)
    assert torch.equal(ops.relu_bwd(dy, y), torch.where(y > 0, dy, torch.zeros_like(dy)))


def test_dynamic_segformer_matches_reference_golden(cuda):
    from gdl_b200.models.segformer import SegFormer
    from oracle import segformer as osf
    gold = torch.load(GOLD / "dynamic_mit_b0_golden.pt")
    sd = osf.init_dynamic_state_dict("mit_b0", 5, seed=4)
    prod = SegFormer("mit_b0", num_classes=5, use_dynamic_encoder=True).cuda().eval()
    prod.load_state_dict(sd)
    for c, case in gold.items():
        x = case["x"].cuda()
        with torch.no_grad():
            got = prod(x)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ac = osf.segformer_forward({k: v.cuda() for k, v in sd.items()}, x, "mit_b0").float()
        ref = case["logits_slice"].cuda()
        ep, ea = _rel(got[:, :, ::4, ::4], ref), _rel(ac[:, :, ::4, ::4], ref)
        print(f"dynamic mit_b0, {c} bands: logits rel err product {ep:.4f}, autocast oracle {ea:.4f}")
        assert ep < max(2.5 * ea, 5e-3)


def test_dynamic_segformer_train_step_parity(cuda):
    from gdl_b200.models.segformer import SegFormer
    from oracle import segformer as osf
    name, cin, k, hw, b = "mit_b0", 4, 5, 128, 4
    torch.manual_seed(0)
    prod = SegFormer(name, num_classes=k, use_dynamic_encoder=True).cuda()
    with torch.no_grad():
        for _, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, cin, hw, hw, generator=g).cuda()
    t = torch.randint(0, k, (b, hw, hw), generator=g).cuda()
    sd = _oracle_sd(prod)
    ref = osf.segformer_forward(sd, x, name, training=True)
    F.cross_entropy(ref, t).backward()
    sd_ac = _oracle_sd(prod)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = osf.segformer_forward(sd_ac, x, name, training=True)
    F.cross_entropy(ac.float(), t).backward()
    prod.train()
    logits = prod(x)
    F.cross_entropy(logits, t).backward()
    e_prod, e_ac = _rel(logits, ref), _rel(ac, ref)
    print(f"dynamic {name}: logits rel err product {e_prod:.4f}, autocast reference {e_ac:.4f}")
    assert e_prod < max(2.5 * e_ac, 5e-3)
    for n, p in prod.named_parameters():
        want = sd[n].grad
        if want is None or want.abs().max() < 1e-9:
            continue
        ep, ea = _rel(p.grad, want), _rel(sd_ac[n].grad, want)
        assert ep < max(3.0 * ea, 2e-2), f"{n}: product {ep:.4f} vs autocast {ea:.4f}"
