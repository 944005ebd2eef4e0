"""GPU: the train-mode stochastic layers of the reference (timm DropPath in MiT, nn.Dropout2d in the SegFormer decoder
and the FCN aux head) with the random draws supplied to product and oracle alike — `gdl_dropout2d_apply` against torch,
and a SegFormer-B0 train step under the same tolerance rule as tests/test_segformer_gpu.py.
(Sorts last on purpose: written after the round's GPU budget was spent; the host logic is pinned on CPU in float64 by
tests/test_engine_host_logic_cpu.py::test_segformer_stochastic_layers_with_supplied_draws.)"""
import pytest
import torch
import torch.nn.functional as F

from test_segformer_gpu import _oracle_sd, _rel, _setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_dropout2d_kernel(cuda, dtype):
    from gdl_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3, 17, 9, 64, generator=g, device="cuda").to(dtype)
    m = ((torch.rand(3, 64, generator=g, device="cuda") < 0.9).float() / 0.9).contiguous()
    y = ops.dropout2d_apply(x, m)
    assert torch.equal(y, (x.float() * m.view(3, 1, 1, 64)).to(dtype))
    wide = torch.randn(3, 17, 9, 96, generator=g, device="cuda").to(dtype)  # channel slice of a wider buffer
    y2 = ops.dropout2d_apply(wide[..., 16:80], m)
    assert torch.equal(y2, (wide[..., 16:80].float() * m.view(3, 1, 1, 64)).to(dtype))
    with pytest.raises(ValueError):
        ops.dropout2d_apply(x, m[:2])


def test_segformer_train_step_with_supplied_draws(cuda):
    from gdl_b200.models.segformer import MIT_CFG
    from oracle import segformer as osf
    name, cin, k, hw, b = "mit_b0", 3, 5, 128, 4
    prod = _setup(name, cin, k)
    nblk = sum(MIT_CFG[name][2])
    g = torch.Generator().manual_seed(5)
    masks = []
    for i in range(nblk):
        keep = 1.0 - 0.1 * i / (nblk - 1)
        masks.append(tuple(((torch.rand(b, generator=g) < keep).float() / keep).cuda() for _ in range(2)))
    masks[2] = (torch.tensor([0.0, 1.25, 1.25, 0.0]).cuda(), torch.tensor([1.25, 0.0, 1.25, 1.25]).cuda())
    dmask = ((torch.rand(b, MIT_CFG[name][3], generator=g) < 0.9).float() / 0.9).cuda()
    prod.drop_path_masks, prod.dropout_mask = masks, dmask
    x = torch.randn(b, cin, hw, hw, generator=g).cuda()
    t = torch.randint(0, k, (b, hw, hw), generator=g).cuda()
    sd = _oracle_sd(prod)
    ref = osf.segformer_forward(sd, x, name, training=True, drop_path=masks, dropout_mask=dmask)
    F.cross_entropy(ref, t).backward()
    sd_ac = _oracle_sd(prod)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = osf.segformer_forward(sd_ac, x, name, training=True, drop_path=masks, dropout_mask=dmask)
    F.cross_entropy(ac.float(), t).backward()
    prod.train()
    logits = prod(x)
    F.cross_entropy(logits, t).backward()
    e_prod, e_ac = _rel(logits, ref), _rel(ac, ref)
    print(f"segformer with DropPath / Dropout2d draws: logits rel err product {e_prod:.4f}, autocast reference {e_ac:.4f}")
    assert e_prod < max(2.5 * e_ac, 5e-3)
    for n, p in prod.named_parameters():
        want = sd[n].grad
        if want.abs().max() < 1e-9:
            continue
        ep, ea = _rel(p.grad, want), _rel(sd_ac[n].grad, want)
        assert ep < max(3.0 * ea, 2e-2), f"{n}: product {ep:.4f} vs autocast {ea:.4f}"
    # rates set, no supplied draws: stochastic in train mode, deterministic in eval mode
    prod.drop_path_masks = prod.dropout_mask = None
    prod.drop_path_rates = [0.3] * nblk
    prod.dropout_ratio = 0.3
    with torch.no_grad():
        prod.eval()
        e1, e2 = prod(x), prod(x)
    assert torch.equal(e1, e2)
