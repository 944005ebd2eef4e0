"""`SegmentationUnetPlus` on the B200 kernels — same constructor keywords, hooks and batch contract
as geo_deep_learning/tasks_with_models/segmentation_unetplus.py:34-248, so the YAML
`class_path: tasks_with_models.segmentation_unetplus.SegmentationUnetPlus` can be pointed here.

What changes underneath: `configure_model` builds gdl_b200.models.unetpp.UnetPlusPlus instead of
smp.UnetPlusPlus (identical state_dict), eval post-processing and the MeanIoU counts use the argmax kernel, and the
kornia augmentation of `on_before_batch_transfer` runs on the device in `on_after_batch_transfer` (tasks/_hooks.py).
Lightning is optional at import time: without it the class derives from a minimal stand-in that
offers `log`/`log_dict`/`save_hyperparameters` no-ops so the step methods stay callable.
"""
from __future__ import annotations

from typing import Any, Callable

import torch
from torch import Tensor

from .. import ops
from ..models.unetpp import UnetPlusPlus

try:  # pragma: no cover - depends on the environment
    from lightning.pytorch import LightningModule as _Base
except Exception:  # noqa: BLE001
    class _Base(torch.nn.Module):
        """Stand-in with the few LightningModule members the step methods touch."""

        def __init__(self) -> None:
            super().__init__()
            self.hparams: dict[str, Any] = {}
            self.logged: dict[str, Any] = {}

        def save_hyperparameters(self, *args: Any, **kwargs: Any) -> None:  # noqa: ARG002
            return None

        def log(self, name: str, value: Any, **kwargs: Any) -> None:  # noqa: ARG002
            self.logged[name] = value

        def log_dict(self, d: dict[str, Any], **kwargs: Any) -> None:  # noqa: ARG002
            self.logged.update(d)


def _strip_model_prefix(sd: dict[str, Tensor]) -> dict[str, Tensor]:
    return {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}


from . import _common  # noqa: E402
from ._hooks import GpuSideHooks  # noqa: E402  (after _Base: the stand-in must exist first)


class SegmentationUnetPlus(GpuSideHooks, _Base):
    def __init__(self, encoder: str, image_size: tuple[int, int], in_channels: int, num_classes: int,
                 max_samples: int, loss: Callable, optimizer: Callable = torch.optim.Adam,
                 scheduler: Callable | None = None, scheduler_config: dict[str, Any] | None = None,
                 weights: str | None = None, class_labels: list[str] | None = None,
                 class_colors: list[str] | None = None, weights_from_checkpoint_path: str | None = None,
                 compute_dtype: torch.dtype = torch.bfloat16, **kwargs: object) -> None:
        super().__init__()
        self.save_hyperparameters(ignore=["loss"])
        self.encoder, self.in_channels, self.num_classes = encoder, in_channels, num_classes
        self.image_size, self.max_samples = image_size, max_samples
        self.loss, self.optimizer, self.scheduler = loss, optimizer, scheduler
        self.scheduler_config = scheduler_config or {"interval": "epoch"}
        self.weights, self.weights_from_checkpoint_path = weights, weights_from_checkpoint_path
        self.class_colors = class_colors
        self.threshold = 0.5
        k = 2 if num_classes == 1 else num_classes
        self.labels = [str(i) for i in range(k)] if class_labels is None else class_labels
        self.compute_dtype = compute_dtype
        self.model: UnetPlusPlus | None = None

    # -- model ---------------------------------------------------------------------------------
    def configure_model(self) -> None:
        if self.model is not None:
            return
        _common.check_pretrained_request(self.weights, "SegmentationUnetPlus")  # smp downloads `encoder_weights` here (:126-131)
        self.model = UnetPlusPlus(encoder_name=self.encoder, in_channels=self.in_channels, encoder_weights=None,
                                  classes=self.num_classes, compute_dtype=self.compute_dtype)
        if self.weights_from_checkpoint_path:
            _common.load_weights_from_checkpoint(self.model, self.weights_from_checkpoint_path,
                                                 _common.hparam(self, "load_parts"),
                                                 trust_pickle=bool(_common.hparam(self, "trust_checkpoint_pickle", False)))

    def configure_optimizers(self):
        """segmentation_unetplus.py:146-205 (shared with the SegFormer / DOFA mirrors: tasks/_common.py)"""
        return _common.configure_optimizers(self)

    def forward(self, image: Tensor) -> Tensor:
        return self.model(image)

    # -- steps ---------------------------------------------------------------------------------
    def training_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        x, y = batch["image"], batch["mask"]
        loss = self.loss(self(x), y)  # the reference passes the mask as is (segmentation_unetplus.py:229-234)
        self.log("train_loss", loss, batch_size=x.shape[0], prog_bar=True, logger=True, on_step=False,
                 on_epoch=True, sync_dist=True, rank_zero_only=True)
        return loss

    def _predict(self, y_hat: Tensor) -> Tensor:
        """softmax(dim=1).argmax(dim=1) | sigmoid > threshold, on the argmax kernel (bit-exact)."""
        nhwc = y_hat.permute(0, 2, 3, 1)
        if nhwc.dtype != torch.float32 or not nhwc.is_contiguous():
            nhwc = nhwc.float().contiguous()
        return ops.argmax_classes(nhwc, self.threshold)

    def validation_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        x, y = batch["image"], batch["mask"]
        y_hat = self(x)
        self.log("val_loss", self.loss(y_hat, y), batch_size=x.shape[0], prog_bar=True, logger=True,
                 on_step=False, on_epoch=True, sync_dist=True, rank_zero_only=True)
        return self._predict(y_hat)

    def test_step(self, batch: dict[str, Any], batch_idx: int) -> None:  # noqa: ARG002
        x, y = batch["image"], batch["mask"]
        y_hat = self(x)
        metrics: dict[str, Any] = {"test_loss": self.loss(y_hat, y)}
        target = y[:, 0] if y.dim() == 4 else y  # the reference squeezes only for the metric (:290)
        _, iou = self._predict_and_score(y_hat, target if target.dtype in (torch.int64, torch.uint8) else target.long())
        metrics.update(iou)
        self.log_dict(metrics, batch_size=x.shape[0], prog_bar=False, logger=True, on_step=False,
                      rank_zero_only=True)
