"""CPU: the Lightning-facing task mirrors (gdl_b200/tasks) with the kernels emulated — constructor keywords, the
training / validation / test step contracts of the reference tasks
(geo_deep_learning/tasks_with_models/segmentation_{unetplus,segformer,dofa}.py), the device-side augmentation hook
and the MeanIoU of test_step against the oracle restatements."""
import pytest
import torch
import torch.nn.functional as F

import cpu_kernel_emulation as emu
from oracle import augment as oaug
from oracle import metrics as omet


def _batch(n, c, hw, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"image": torch.randn(n, c, hw, hw, generator=g),
            "mask": torch.randint(0, k, (n, 1, hw, hw), generator=g)}


def _check_iou(logged, logits, target, labels):
    pred = logits.argmax(1)
    s, c = omet.mean_iou_update(pred, target, len(labels))
    want = omet.mean_iou_compute(s, c)
    for i, name in enumerate(labels):
        assert abs(float(logged[f"meaniou_{name}"]) - float(want[i])) < 1e-6


def test_unetplus_task_steps(monkeypatch):
    from gdl_b200.tasks.segmentation_unetplus import SegmentationUnetPlus
    emu.install(monkeypatch)
    torch.manual_seed(0)
    task = SegmentationUnetPlus("resnet18", (32, 32), 3, 4, max_samples=2, loss=torch.nn.CrossEntropyLoss(),
                                class_labels=["bg", "a", "b", "c"], compute_dtype=torch.float32)
    task.configure_model()
    task.configure_model()  # idempotent, as Lightning may call it twice
    assert sorted(task.state_dict())[0].startswith("model.")
    batch = _batch(2, 3, 32, 4)
    # the reference's UNet++ task hands the mask to the loss as is (:229-234): CrossEntropyLoss wants (N,H,W) int64
    batch["mask"] = batch["mask"][:, 0]
    task.train()
    loss = task.training_step(batch, 0)
    loss.backward()
    assert torch.isfinite(loss) and task.logged["train_loss"] is loss
    assert all(p.grad is not None for p in task.model.parameters() if p.requires_grad)
    task.eval()
    with torch.no_grad():
        pred = task.validation_step(batch, 0)
        logits = task(batch["image"])
        assert pred.dtype == torch.int64 and torch.equal(pred, logits.argmax(1))
        task.test_step(batch, 0)
    assert abs(float(task.logged["test_loss"]) - float(F.cross_entropy(logits, batch["mask"]))) < 1e-6
    _check_iou(task.logged, logits, batch["mask"], task.labels)
    opt = task.configure_optimizers()
    assert isinstance(opt[0], torch.optim.Adam)


def test_segformer_task_steps_and_device_side_augmentation(monkeypatch):
    from gdl_b200.tasks.segmentation_segformer import SegmentationSegformer
    emu.install(monkeypatch)
    torch.manual_seed(0)
    task = SegmentationSegformer("mit_b0", image_size=(64, 64), in_channels=4, num_classes=3, max_samples=2,
                                 loss=torch.nn.CrossEntropyLoss(), compute_dtype=torch.float32)
    task.configure_model()
    batch = _batch(3, 4, 64, 3, seed=2)
    # eval: the hook leaves the batch alone (the reference only augments when trainer.training)
    task.eval()
    same = task.on_after_batch_transfer(dict(batch), 0)
    assert same["image"] is batch["image"] and same["mask"] is batch["mask"]
    # train: one drawn operation applied per sample; image and mask move together, shapes / dtypes are kept
    task.train()
    torch.manual_seed(5)
    out = task.on_after_batch_transfer(dict(batch), 0)
    assert out["image"].shape == batch["image"].shape and out["mask"].shape == batch["mask"].shape
    assert out["mask"].dtype == torch.int64 and out["image"].dtype == torch.float32
    torch.manual_seed(5)
    params = task._augmenter.__class__((64, 64)).sample(3)  # the same draws from the same global RNG state
    want_i, want_m = oaug.apply_params(batch["image"], batch["mask"][:, 0], params)
    assert torch.equal(out["mask"][:, 0], want_m) and (out["image"] - want_i).abs().max() < 1e-5
    loss = task.training_step(out, 0)
    loss.backward()
    assert torch.isfinite(loss)
    task.eval()
    with torch.no_grad():
        task.test_step(batch, 0)
        logits = task(batch["image"])
    _check_iou(task.logged, logits, batch["mask"][:, 0], task.labels)
    task.gpu_augment = False
    task.train()
    assert task.on_after_batch_transfer(dict(batch), 0)["image"] is batch["image"]


def test_dofa_task_steps(monkeypatch):
    from gdl_b200.tasks.segmentation_dofa import SegmentationDOFA
    emu.install(monkeypatch)
    torch.manual_seed(0)
    task = SegmentationDOFA("dofa_base", pretrained=False, image_size=(56, 56), num_classes=3, max_samples=2,
                            loss=torch.nn.CrossEntropyLoss(), freeze_layers=["encoder"], compute_dtype=torch.float32)
    task.configure_model()
    batch = _batch(2, 3, 56, 3, seed=4)
    batch["wavelengths"] = torch.tensor([0.665, 0.56, 0.49])
    task.train()
    loss = task.training_step(batch, 0)
    loss.backward()
    assert torch.isfinite(loss)
    assert all(p.grad is None for n, p in task.model.named_parameters() if n.startswith("encoder."))
    task.eval()
    with torch.no_grad():
        out = task(batch["image"], batch["wavelengths"])
        task.test_step(batch, 0)
    y = batch["mask"][:, 0]
    want = F.cross_entropy(out.out, y) + 0.4 * F.cross_entropy(out.aux, y)
    assert abs(float(task.logged["test_loss"]) - float(want)) < 1e-5
    _check_iou(task.logged, out.out, y, task.labels)


def test_quickstart_construction_and_one_batch_each(monkeypatch):
    """The reference's own integration tests (tests/test_notebooks_00quickstart.py:52-71,101-118) against the task mirror:
    the same constructor arguments, and — Lightning is not installed here — a hand-rolled `fast_dev_run`: one batch of
    train (step + backward + optimizer + scheduler), validation and test from the reference's RandomDataset recipe."""
    from torch.utils.data import DataLoader, Dataset

    from gdl_b200.tasks.segmentation_unetplus import SegmentationUnetPlus
    emu.install(monkeypatch)

    class RandomDataset(Dataset):
        def __len__(self):
            return 4

        def __getitem__(self, idx):
            return {"image": torch.rand(3, 32, 32), "mask": torch.zeros(32, 32, dtype=torch.long)}

    model = SegmentationUnetPlus(encoder="resnet34", in_channels=3, num_classes=2, image_size=(64, 64), max_samples=1,
                                 loss=torch.nn.CrossEntropyLoss(), optimizer=lambda params: torch.optim.Adam(params, lr=1e-3),
                                 scheduler=torch.optim.lr_scheduler.StepLR, scheduler_config={"step_size": 1, "gamma": 0.1},
                                 class_labels=["background", "buildings"], class_colors=["#000000", "#FF0000"])
    assert hasattr(model, "training_step") and isinstance(model, torch.nn.Module)
    torch.manual_seed(0)
    model = SegmentationUnetPlus(encoder="resnet34", in_channels=3, num_classes=2, image_size=(32, 32), max_samples=1,
                                 loss=torch.nn.CrossEntropyLoss(), optimizer=lambda params: torch.optim.Adam(params, lr=1e-3),
                                 scheduler=lambda opt: torch.optim.lr_scheduler.StepLR(opt, step_size=1), scheduler_config={},
                                 class_labels=["background", "buildings"], class_colors=["#000000", "#FF0000"],
                                 compute_dtype=torch.float32)
    model.configure_model()
    (opt,), (sched_cfg,) = model.configure_optimizers()
    assert sched_cfg["interval"] == "epoch" and isinstance(sched_cfg["scheduler"], torch.optim.lr_scheduler.StepLR)
    loader = DataLoader(RandomDataset(), batch_size=2)
    model.train()
    batch = model.on_after_batch_transfer(next(iter(loader)), 0)
    w0 = model.model.segmentation_head[0].weight.detach().clone()
    loss = model.training_step(batch, 0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    sched_cfg["scheduler"].step()
    assert torch.isfinite(loss) and not torch.equal(model.model.segmentation_head[0].weight, w0)
    assert abs(opt.param_groups[0]["lr"] - 1e-4) < 1e-12
    model.eval()
    with torch.no_grad():
        pred = model.validation_step(next(iter(loader)), 0)
        model.test_step(next(iter(loader)), 0)
    assert pred.shape == (2, 32, 32) and {"test_loss", "meaniou_background", "meaniou_buildings"} <= set(model.logged)


def test_unetplus_onecycle_from_cli_hyperparameters(monkeypatch):
    """The LightningCLI branch of configure_optimizers (segmentation_unetplus.py:160-200): OneCycleLR sized from the trainer."""
    from types import SimpleNamespace

    from gdl_b200.tasks.segmentation_unetplus import SegmentationUnetPlus
    task = SegmentationUnetPlus("resnet18", (32, 32), 3, 2, max_samples=1, loss=torch.nn.CrossEntropyLoss(),
                                optimizer=lambda p: torch.optim.Adam(p, lr=1e-3), compute_dtype=torch.float32)
    task.configure_model()
    cfg = {"class_path": "torch.optim.lr_scheduler.OneCycleLR", "init_args": {"max_lr": 0.01, "total_steps": 77}}
    task.hparams = {"scheduler": cfg}
    task.trainer = SimpleNamespace(estimated_stepping_batches=120, datamodule=None, accumulate_grad_batches=1, max_epochs=3)
    _, (s1,) = task.configure_optimizers()
    assert isinstance(s1["scheduler"], torch.optim.lr_scheduler.OneCycleLR) and s1["scheduler"].total_steps == 120
    task.trainer = SimpleNamespace(estimated_stepping_batches=-1, datamodule=SimpleNamespace(epoch_size=100, batch_size=8),
                                   accumulate_grad_batches=2, max_epochs=3)
    _, (s2,) = task.configure_optimizers()
    assert s2["scheduler"].total_steps == (7 + 14) * 3  # ceil(100 / 16) = 7 steps + 7 * 2 buffer, 3 epochs
    task.trainer = SimpleNamespace(estimated_stepping_batches=-1, datamodule=None, accumulate_grad_batches=1, max_epochs=3)
    _, (s3,) = task.configure_optimizers()
    assert s3["scheduler"].total_steps == 77


def test_segformer_and_dofa_mirrors_share_checkpoint_and_scheduler_logic(monkeypatch, tmp_path):
    """the SegFormer / DOFA mirrors must (a) honour `load_parts` (utils/models.py:31-66: filter + strict=False),
    (b) size OneCycleLR from the trainer like the reference's configure_optimizers (segmentation_segformer.py:150-199),
    (c) refuse a pretrained-weights request loudly instead of training from random initialisation, (d) load checkpoints
    with torch.load's default weights_only=True."""
    from types import SimpleNamespace

    from gdl_b200.tasks.segmentation_segformer import SegmentationSegformer
    from gdl_b200.tasks.segmentation_unetplus import SegmentationUnetPlus
    emu.install(monkeypatch)

    def make(**kw):
        return SegmentationSegformer("mit_b0", image_size=(64, 64), in_channels=3, num_classes=3, max_samples=1,
                                     loss=torch.nn.CrossEntropyLoss(), optimizer=lambda p: torch.optim.Adam(p, lr=1e-3),
                                     compute_dtype=torch.float32, **kw)
    torch.manual_seed(0)
    src = make()
    src.configure_model()
    ckpt = tmp_path / "ckpt.pt"
    torch.save({"state_dict": {f"model.{k}": v for k, v in src.model.state_dict().items()}}, ckpt)
    # (a) only the encoder is taken from the checkpoint; the decoder keeps its own initialisation
    torch.manual_seed(1)
    dst = make(weights_from_checkpoint_path=str(ckpt))
    dst.hparams = {"load_parts": ["encoder"]}
    dst.configure_model()
    a, b = src.model.state_dict(), dst.model.state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a if k.startswith("encoder."))
    assert any(not torch.equal(a[k], b[k]) for k in a if k.startswith("decoder.") and a[k].is_floating_point() and a[k].dim() > 1)
    # full load: everything
    full = make(weights_from_checkpoint_path=str(ckpt))
    full.configure_model()
    assert all(torch.equal(a[k], v) for k, v in full.model.state_dict().items())
    # (d) a checkpoint with arbitrary pickled objects is rejected unless explicitly trusted
    bad = tmp_path / "bad.pt"
    torch.save({"state_dict": a, "callback": SimpleNamespace(x=1)}, bad)
    risky = make(weights_from_checkpoint_path=str(bad))
    with pytest.raises(Exception):
        risky.configure_model()
    # (b) OneCycleLR horizon from the trainer
    cfg = {"class_path": "torch.optim.lr_scheduler.OneCycleLR", "init_args": {"max_lr": 0.01, "total_steps": 55}}
    full.hparams = {"scheduler": cfg}
    full.trainer = SimpleNamespace(estimated_stepping_batches=90, datamodule=None, accumulate_grad_batches=1, max_epochs=2)
    _, (s1,) = full.configure_optimizers()
    assert isinstance(s1["scheduler"], torch.optim.lr_scheduler.OneCycleLR) and s1["scheduler"].total_steps == 90
    full.trainer = SimpleNamespace(estimated_stepping_batches=-1, datamodule=None, accumulate_grad_batches=1, max_epochs=2)
    _, (s2,) = full.configure_optimizers()
    assert s2["scheduler"].total_steps == 55
    # (c) `weights: imagenet` (the shipped YAMLs) must not be dropped silently
    for task in (make(weights="imagenet"),
                 SegmentationUnetPlus("resnet18", (32, 32), 3, 2, max_samples=1, loss=torch.nn.CrossEntropyLoss(),
                                      weights="imagenet", compute_dtype=torch.float32)):
        with pytest.raises(ValueError, match="pretrained"):
            task.configure_model()
