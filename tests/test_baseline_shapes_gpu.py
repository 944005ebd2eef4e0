"""GPU: parity at the shapes BASELINE.json quotes — 512 x 512 tiles — not only at the 64..128-pixel tiles of the model
tests (tile scheduling, the wgrad pixel split / partial tiles, the row-streaming kernels' block
geometry and 32-bit index arithmetic differ at full size).  Same bars as the model tests: deviation from the fp32
oracle <= 2.5x (logits, loss) / 3x (every parameter gradient) the deviation of the reference stack itself under
torch.autocast on the same inputs; argmax agreement with the fp32 oracle >= the autocast reference's - 0.5 %.
The numbers are printed (pytest -rA / -s) so the run log is the parity report.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def test_unetpp_r50_4band_512_train_step(cuda):
    """BASELINE configs[1]: UNet++-ResNet50, 4-band 512 x 512, 5 classes, bf16 (batch 4 of the 32)."""
    from test_unetpp_gpu import test_train_step_parity
    test_train_step_parity(cuda, "resnet50", 4, 5, 512, torch.bfloat16)


def test_unetpp_r18_3band_256_train_step(cuda):
    """BASELINE configs[0]: UNet++-ResNet18, 3-band 256 x 256, 5 classes, batch 4 (the reference's CPU-runnable case)."""
    from test_unetpp_gpu import test_train_step_parity
    test_train_step_parity(cuda, "resnet18", 3, 5, 256, torch.bfloat16)


def test_segformer_b2_3band_512_train_step(cuda):
    """BASELINE configs[2]: SegFormer-B2 (MixTransformer), 3-band 512 x 512, bf16 (batch 4 of the 16 per GPU)."""
    from test_segformer_gpu import train_step_parity
    train_step_parity("mit_b2", 3, 5, 512, torch.bfloat16)  # the default decoder route (folded)


def test_segformer_b5_4band_512_inference(cuda):
    """BASELINE configs[4]: SegFormer-B5, 4-band 512 x 512 windows, inference: logits and class masks of one window batch."""
    from oracle import segformer as osf
    from test_segformer_gpu import _setup
    prod = _setup("mit_b5", 4, 5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 512, 512, generator=g).cuda()
    sd = {k: v.detach() for k, v in prod.state_dict().items()}
    prod.eval()
    with torch.no_grad():
        ref = osf.segformer_forward(sd, x, "mit_b5", training=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = osf.segformer_forward(sd, x, "mit_b5", training=False).float()
        out = prod(x)
        cls = prod.predict_classes(x)
    e_prod, e_ac = _rel(out, ref), _rel(ac, ref)
    agree, agree_ac = (cls == ref.argmax(1)).float().mean().item(), (ac.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"[mit_b5 512 eval] logits rel err product {e_prod:.4f}, autocast reference {e_ac:.4f}; "
          f"argmax agreement with the fp32 oracle: product {agree:.5f}, autocast reference {agree_ac:.5f}")
    assert e_prod < max(2.5 * e_ac, 5e-3)
    assert agree >= agree_ac - 0.005
    assert torch.equal(cls, out.argmax(1))  # the fused upsample + argmax head == argmax of the materialised logits


def test_dofa_base_upernet_6band_512_train_step(cuda):
    """BASELINE configs[3]: DOFA-base (frozen) + UperNet, 6 bands at the multi-sensor wavelengths, 512 x 512 (batch 2)."""
    from test_dofa_gpu import test_dofa_segmentation_train_step
    test_dofa_segmentation_train_step(cuda, img=512, bands=6, batch=2,
                                      wavelengths=(0.49, 0.56, 0.665, 0.842, 1.61, 2.19))


def test_fused_trainer_at_full_batch_is_finite_and_reproducible(cuda):
    """configs[1] at its full batch (32 x 4 x 512 x 512) through the fused trainer: two steps from the same state agree bit
    for bit (ordered reductions at full size: 2048-block slot sums, ~450 partial wgrad tiles per layer) and stay finite."""
    from gdl_b200 import ops
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer
    g = torch.Generator().manual_seed(7)
    raw = torch.randint(0, 256, (32, 512, 512, 4), generator=g, dtype=torch.uint8).cuda()
    t = torch.randint(0, 5, (32, 16, 16), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    out = []
    for _ in range(2):
        torch.manual_seed(0)
        m = UnetPlusPlus("resnet50", in_channels=4, classes=5).cuda().train()
        tr = FusedTrainer(m, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-4, mean=[0.5] * 4, std=[0.2] * 4)
        loss = tr.forward_backward(raw, t)
        out.append((loss.item(), tr.gflat.clone()))
        del tr, m
        torch.cuda.empty_cache()
    assert torch.isfinite(out[0][1]).all() and out[0][0] == out[0][0]
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
    print(f"[unetpp_r50 B=32 512] loss {out[0][0]:.5f}, |grad| {out[0][1].norm().item():.4f}")
