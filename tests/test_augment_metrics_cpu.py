"""CPU: the augmentation / metric widening without a GPU.

* the kernel's index and weight arithmetic (transcribed line by line in tests/cpu_kernel_emulation.py::augment_normalize)
  against torch's own flip / rot90 / F.interpolate (oracle/augment.py) — exact for the permutations and the mask, fp32
  rounding for the bilinear crop;
* the host-side draws of gdl_b200.augment.BatchAugmenter (kornia's published sampling rules, restated);
* the MeanIoU accumulation of gdl_b200.metrics against the one-hot restatement of torchmetrics 1.8 (oracle/metrics.py).
"""
import pytest
import torch

import cpu_kernel_emulation as emu
from oracle import augment as oaug
from oracle import metrics as omet
from oracle import tensors as ot


def _batch(n, c, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randint(0, 256, (n, h, w, c), generator=g, dtype=torch.uint8)
    mask = torch.randint(0, 5, (n, h, w), generator=g, dtype=torch.uint8)
    return raw, mask


def _reference(raw, mask, params, mean, std):
    x = ot.standardization(ot.normalization(raw.permute(0, 3, 1, 2).float()), mean.view(-1, 1), std.view(-1, 1))
    return oaug.apply_params(x, mask, params)


@pytest.mark.parametrize("c", [3, 4, 6])
def test_kernel_arithmetic_exact_ops_equal_torch(c):
    n, h, w = 9, 24, 24
    raw, mask = _batch(n, c, h, w)
    mean, std = torch.full((c,), 0.45), torch.full((c,), 0.21)
    params = torch.zeros(n, 6, dtype=torch.int32)
    params[:, 0] = torch.tensor([0, 1, 2, 3, 3, 3, 3, 1, 2])
    params[:, 1] = torch.tensor([0, 0, 0, 1, 2, 3, 0, 0, 0])
    want_i, want_m = _reference(raw, mask, params, mean, std)
    got_i, got_m = emu.augment_normalize(raw, False, mask, params, torch.float32, 0, mean, std, 255.0)
    assert torch.equal(got_m, want_m)
    assert torch.equal(got_i, want_i)  # permutations commute exactly with the per-pixel normalisation
    # 16-bit NHWC output: the same values after one rounding, zero padded to 8 channels
    got16, _ = emu.augment_normalize(raw, False, mask, params, torch.bfloat16, 8, mean, std, 255.0)
    assert got16.shape == (n, h, w, 8) and torch.equal(got16[..., :c], want_i.permute(0, 2, 3, 1).bfloat16())
    assert not got16[..., c:].any()


@pytest.mark.parametrize("hw", [(32, 32), (40, 24)])
def test_kernel_arithmetic_resized_crop_equals_interpolate(hw):
    h, w = hw
    n, c = 12, 4
    raw, mask = _batch(n, c, h, w, seed=3)
    mean, std = torch.tensor([0.4, 0.5, 0.6, 0.3]), torch.tensor([0.2, 0.25, 0.3, 0.15])
    g = torch.Generator().manual_seed(5)
    params = torch.zeros(n, 6, dtype=torch.int32)
    params[:, 0] = oaug.CROP
    for i in range(n):
        ch = int(torch.randint(1, h + 1, (1,), generator=g))
        cw = int(torch.randint(1, w + 1, (1,), generator=g))
        params[i, 2] = int(torch.randint(0, h - ch + 1, (1,), generator=g))
        params[i, 3] = int(torch.randint(0, w - cw + 1, (1,), generator=g))
        params[i, 4], params[i, 5] = ch, cw
    params[0, 2:] = torch.tensor([0, 0, h, w], dtype=torch.int32)      # whole tile: identity resize
    params[1, 2:] = torch.tensor([h - 1, w - 1, 1, 1], dtype=torch.int32)  # one pixel blown up
    want_i, want_m = _reference(raw, mask, params, mean, std)
    got_i, got_m = emu.augment_normalize(raw, False, mask, params, torch.float32, 0, mean, std, 255.0)
    assert torch.equal(got_m, want_m)  # nearest: ATen's floor(dst * in/out) index arithmetic, bit exact
    # bilinear: interpolate-then-normalise (kernel) vs normalise-then-interpolate (reference) differ by fp32 rounding only
    assert (got_i - want_i).abs().max() < 2e-5
    # NCHW float input (the Lightning batch layout) goes through the same arithmetic
    xf = raw.permute(0, 3, 1, 2).float().contiguous()
    got_f, _ = emu.augment_normalize(xf, True, mask, params, torch.float32, 0, mean, std, 255.0)
    assert torch.equal(got_f, got_i)


def test_int64_mask_and_no_mask():
    raw, mask = _batch(3, 3, 16, 16)
    params = torch.tensor([[1, 0, 0, 0, 0, 0], [4, 0, 2, 3, 9, 7], [3, 3, 0, 0, 0, 0]], dtype=torch.int32)
    _, m8 = emu.augment_normalize(raw, False, mask, params, torch.float32)
    _, m64 = emu.augment_normalize(raw, False, mask.long(), params, torch.float32)
    assert m64.dtype == torch.int64 and torch.equal(m64, m8.long())
    img, none = emu.augment_normalize(raw, False, None, params, torch.float32)
    assert none is None and img.shape == (3, 3, 16, 16)


def test_sampler_follows_the_published_rules():
    from gdl_b200 import ops
    from gdl_b200.augment import OPS, BatchAugmenter
    aug = BatchAugmenter((64, 64), generator=torch.Generator().manual_seed(0))
    seen = {name: 0 for name in OPS}
    applied = {name: 0 for name in OPS}
    for _ in range(400):
        p = aug.sample(16)
        assert p.dtype == torch.int32 and p.shape == (16, 6)
        name = aug.last_op
        seen[name] += 1
        ops_used = set(p[:, 0].tolist())
        if name == "hflip":
            assert ops_used <= {ops.AUG_IDENTITY, ops.AUG_HFLIP}
        elif name == "vflip":
            assert ops_used <= {ops.AUG_IDENTITY, ops.AUG_VFLIP}
        elif name == "rot90":
            assert ops_used <= {ops.AUG_IDENTITY, ops.AUG_ROT90}
            r = p[p[:, 0] == ops.AUG_ROT90]
            assert ((r[:, 1] >= 1) & (r[:, 1] <= 3)).all()
        elif name == "resized_crop_zoom_in":
            assert ops_used == {ops.AUG_IDENTITY}  # scale (1, 2): no window fits strictly inside the tile
        else:
            assert ops_used <= {ops.AUG_IDENTITY, ops.AUG_CROP}
            cr = p[p[:, 0] == ops.AUG_CROP].long()
            area = cr[:, 4] * cr[:, 5]
            assert (area >= 0.45 * 64 * 64).all() and (area < 64 * 64).all()
            assert ((cr[:, 4] < 64) & (cr[:, 5] < 64) & (cr[:, 4] > 0) & (cr[:, 5] > 0)).all()
            assert ((cr[:, 2] >= 0) & (cr[:, 2] + cr[:, 4] <= 64) & (cr[:, 3] >= 0) & (cr[:, 3] + cr[:, 5] <= 64)).all()
            ratio = cr[:, 5].float() / cr[:, 4].float()
            assert (ratio > 0.7).all() and (ratio < 1.4).all()
        applied[name] += int((p[:, 0] != ops.AUG_IDENTITY).sum())
    assert all(50 <= v <= 110 for v in seen.values()), seen          # one of five, uniformly (expected 80 each)
    for name in ("hflip", "vflip", "rot90", "resized_crop_zoom_out"):
        frac = applied[name] / (16 * seen[name])
        assert 0.42 < frac < 0.58, (name, frac)                        # each sample with p = 0.5
    # same seed, same draws
    a = BatchAugmenter((64, 64), generator=torch.Generator().manual_seed(7)).sample(8)
    b = BatchAugmenter((64, 64), generator=torch.Generator().manual_seed(7)).sample(8)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        BatchAugmenter((64, 64), p=1.5)


def test_augmenter_call_routes_through_the_kernel_wrapper(monkeypatch):
    from gdl_b200.augment import BatchAugmenter
    emu.install(monkeypatch)
    raw, mask = _batch(6, 4, 32, 32, seed=9)
    mean, std = torch.full((4,), 0.5), torch.full((4,), 0.2)
    aug = BatchAugmenter((32, 32), generator=torch.Generator().manual_seed(2))
    params = aug.sample(6)
    img, m = aug(raw, mask, chw=False, out_dtype=torch.float32, mean=mean, std=std, image_max=255.0, params=params)
    want_i, want_m = _reference(raw, mask, params, mean, std)
    assert torch.equal(m, want_m) and (img - want_i).abs().max() < 2e-5
    with pytest.raises(ValueError):
        aug(raw[:, :16], mask[:, :16], chw=False, out_dtype=torch.float32)


@pytest.mark.parametrize("k", [1, 5])
def test_mean_iou_matches_the_torchmetrics_restatement(monkeypatch, k):
    from gdl_b200.metrics import MeanIoU
    emu.install(monkeypatch)
    kc = 2 if k == 1 else k
    g = torch.Generator().manual_seed(11)
    metric = MeanIoU(kc, labels=[f"c{i}" for i in range(kc)])
    s_tot, c_tot = torch.zeros(kc, dtype=torch.float64), torch.zeros(kc, dtype=torch.int64)
    for b in range(3):
        logits = torch.randn(4, 16, 16, k, generator=g)
        target = torch.randint(0, kc, (4, 16, 16), generator=g)
        if b == 1:
            target[0] = 0                       # a sample in which most classes are absent from the target
            logits[0, ..., 0] += 50.0           # ... and from the prediction: their unions are empty there
        cls = metric.update(logits, target)
        want_cls = logits.argmax(3) if k > 1 else (logits[..., 0].sigmoid() > 0.5).long()
        assert torch.equal(cls, want_cls)
        conf = emu.argmax_confusion(logits, target)[1]
        assert torch.equal(conf, omet.confusion_per_sample(want_cls, target, kc))
        s, c = omet.mean_iou_update(want_cls, target, kc)
        s_tot += s
        c_tot += c
    want = omet.mean_iou_compute(s_tot, c_tot)
    got = metric.compute()
    assert list(got) == [f"meaniou_c{i}" for i in range(kc)]
    assert torch.allclose(torch.stack(list(got.values())).double(), want, atol=1e-6)
    metric.reset()
    with pytest.raises(RuntimeError):
        metric.compute()


def test_confusion_skips_ignored_and_out_of_range_targets():
    logits = torch.randn(2, 8, 8, 3)
    target = torch.randint(0, 3, (2, 8, 8))
    target[0, :2] = 255
    target[1, 0] = 7
    _, conf = emu.argmax_confusion(logits, target, ignore_index=255)
    assert conf.sum() == 2 * 64 - 16 - 8


def test_mean_iou_restatement_agrees_with_sklearn_per_sample():
    """Independent cross-check of oracle.metrics (torchmetrics itself is un-vendored: parity stays unpinned): its per-sample,
    per-class score is the Jaccard index, which scikit-learn computes from its own confusion-matrix code; the metric is the
    mean of those scores over the samples in which the class occurs (in the prediction or the target)."""
    skm = pytest.importorskip("sklearn.metrics")
    k, n = 5, 6
    g = torch.Generator().manual_seed(11)
    target = torch.randint(0, k, (n, 24, 24), generator=g)
    pred = torch.where(torch.rand(n, 24, 24, generator=g) < 0.7, target, torch.randint(0, k, (n, 24, 24), generator=g))
    target[0][target[0] == 3] = 1  # class 3 absent from sample 0's target ...
    pred[0][pred[0] == 3] = 1      # ... and prediction: not a valid sample for that class
    score_sum, valid = omet.mean_iou_update(pred, target, k)
    want_sum, want_valid = torch.zeros(k, dtype=torch.float64), torch.zeros(k, dtype=torch.long)
    for i in range(n):
        t_, p_ = target[i].reshape(-1).numpy(), pred[i].reshape(-1).numpy()
        j = skm.jaccard_score(t_, p_, labels=list(range(k)), average=None, zero_division=0.0)
        present = torch.tensor([(t_ == c).any() or (p_ == c).any() for c in range(k)])
        want_sum += torch.tensor(j, dtype=torch.float64) * present
        want_valid += present.long()
    assert want_valid[3] == n - 1
    assert torch.equal(valid, want_valid)
    assert torch.allclose(score_sum, want_sum, atol=1e-12)
    assert torch.allclose(omet.mean_iou_compute(score_sum, valid), want_sum / want_valid, atol=1e-12)
