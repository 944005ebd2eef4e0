"""GPU: the WebDataset shard feeder staging raw uint8 batches through pinned memory onto the device, normalised there by
the kernels — against the REFERENCE's own `_process_sample` outputs (tests/golden/wds_golden.pt).
(Sorts last on purpose: written after the round's GPU budget was spent; the CPU suite covers the same logic with the
kernels emulated.)"""
import pytest
import torch

from test_wds_feeder_cpu import KEYS, _dataset, _gold

pytestmark = pytest.mark.gpu


def test_feeder_to_device_matches_reference_golden(cuda, tmp_path):
    from gdl_b200 import ops
    from gdl_b200 import wds_feeder as wf
    sensor, samples = _dataset(tmp_path)
    stats = wf.load_normalization_stats(str(tmp_path / "stats.json"), sensor)
    paths, count = wf.create_shard_split_paths(str(tmp_path / "manifest.json"), "val")
    feeder = wf.ShardFeeder(sensor, paths, stats, model_type="dofa", split="val", batch_size=4, device="cuda",
                            wavelength_keys=KEYS, normalize=True, rank=0, world_size=1)
    batches = list(feeder)
    assert sum(b["image_u8"].shape[0] for b in batches) == count
    g = _gold()
    b0 = batches[0]
    assert b0["image_u8"].is_cuda and b0["mask"].is_cuda and b0["image_u8"].dtype == torch.uint8
    for j, ref in enumerate(g["outputs"]["dofa"]):
        assert torch.equal(b0["mask"][j].cpu(), ref["mask"])
        assert torch.equal(b0["image"][j].cpu(), ref["image"])      # same fp32 divisions, same order: bit equal
        assert torch.equal(b0["wavelengths"][j].cpu(), ref["wavelengths"])
    # the 16-bit NHWC operand of the stem straight from the raw CHW batch == one rounding of the reference image
    x16 = ops.normalize_to_nhwc(b0["image_u8"], True, torch.bfloat16, 8, b0["mean"], b0["std"], 255.0)
    for j, ref in enumerate(g["outputs"]["dofa"]):
        assert torch.equal(x16[j, :, :, :4].cpu(), ref["image"].permute(1, 2, 0).bfloat16())


def test_trainer_steps_from_the_feeder(cuda, tmp_path):
    from gdl_b200 import ops
    from gdl_b200 import wds_feeder as wf
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer
    sensor, _ = _dataset(tmp_path, n_shards=2, per_shard=4, c=4, hw=64, seed=5)
    stats = wf.load_normalization_stats(str(tmp_path / "stats.json"), sensor)
    paths, _ = wf.create_shard_split_paths(str(tmp_path / "manifest.json"), "trn")
    torch.manual_seed(0)
    model = UnetPlusPlus("resnet18", in_channels=4, classes=5).cuda().train()
    tr = FusedTrainer(model, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=stats["mean"].tolist(),
                      std=stats["std"].tolist(), input_chw=True)
    losses = []
    for _ in range(3):
        for b in wf.ShardFeeder(sensor, paths, stats, split="trn", batch_size=4, device="cuda", rank=0, world_size=1,
                                shuffle_buffer=4, mask_dtype=torch.uint8):
            losses.append(float(tr.step(b["image_u8"], b["mask"][:, 0].contiguous())))
    assert len(losses) == 6 and all(l == l for l in losses) and min(losses[3:]) < losses[0]
