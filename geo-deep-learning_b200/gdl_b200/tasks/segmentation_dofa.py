"""`SegmentationDOFA` on the B200 kernels — constructor keywords, hooks and batch contract of
geo_deep_learning/tasks_with_models/segmentation_dofa.py:33-345: `configure_model` builds
gdl_b200.models.dofa.DOFASegmentationModel (same state_dict as the reference's), the batch carries
`image`, `mask` and `wavelengths` (:219-221), the loss is `loss(out) + 0.4 * loss(aux)` (:226-228, :264-266) and
the eval post-processing (`softmax(dim=1).argmax(dim=1)` / `sigmoid > 0.5`, :280-283) runs on the argmax kernel."""
from __future__ import annotations

from typing import Any, Callable

import torch
from torch import Tensor

from ..models.dofa import DOFASegmentationModel, SegmentationOutput
from . import _common
from ._hooks import GpuSideHooks
from .segmentation_segformer import SegmentationSegformer
from .segmentation_unetplus import _Base, _strip_model_prefix


class SegmentationDOFA(GpuSideHooks, _Base):
    def __init__(self, encoder: str, *, pretrained: bool, image_size: tuple[int, int], num_classes: int,
                 max_samples: int, loss: Callable, optimizer: Callable = torch.optim.Adam,
                 scheduler: Callable | None = None, scheduler_config: dict[str, Any] | None = None,
                 freeze_layers: list[str] | None = None, class_labels: list[str] | None = None,
                 class_colors: list[str] | None = None, weights_from_checkpoint_path: str | None = None,
                 compute_dtype: torch.dtype = torch.bfloat16, **kwargs: object) -> None:
        super().__init__()
        self.save_hyperparameters(ignore=["loss"])
        self.encoder, self.pretrained, self.image_size = encoder, pretrained, tuple(image_size)
        self.freeze_layers, self.weights_from_checkpoint_path = freeze_layers, weights_from_checkpoint_path
        self.optimizer, self.scheduler = optimizer, scheduler
        self.scheduler_config = scheduler_config or {"interval": "epoch"}
        self.class_colors, self.max_samples, self.num_classes = class_colors, max_samples, num_classes
        self.threshold = 0.5
        self.loss = loss
        k = 2 if num_classes == 1 else num_classes
        self.labels = [str(i) for i in range(k)] if class_labels is None else class_labels
        self.compute_dtype = compute_dtype
        self.model: DOFASegmentationModel | None = None

    def configure_model(self) -> None:
        if self.model is not None:
            return
        self.model = DOFASegmentationModel(self.encoder, self.image_size, self.freeze_layers, self.num_classes,
                                           pretrained=self.pretrained, compute_dtype=self.compute_dtype,
                                           aux_dropout_ratio=0.1)  # FCNHead's Dropout2d default (fcn_head.py:24)
        if self.weights_from_checkpoint_path:
            _common.load_weights_from_checkpoint(self.model, self.weights_from_checkpoint_path,
                                                 _common.hparam(self, "load_parts"),
                                                 trust_pickle=bool(_common.hparam(self, "trust_checkpoint_pickle", False)))

    configure_optimizers = SegmentationSegformer.configure_optimizers
    _predict = SegmentationSegformer._predict

    def forward(self, image: Tensor, wavelengths: Tensor) -> SegmentationOutput:
        return self.model(image, wavelengths)

    def _loss(self, batch: dict[str, Any]):
        x, y = batch["image"], batch["mask"].squeeze(1).long()
        outputs = self(x, batch["wavelengths"])
        return outputs, self.loss(outputs.out, y) + 0.4 * self.loss(outputs.aux, y), x.shape[0]

    def training_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        _, loss, bs = self._loss(batch)
        self.log("train_loss", loss, batch_size=bs, prog_bar=True, logger=True, on_step=False, on_epoch=True,
                 sync_dist=True, rank_zero_only=True)
        return loss

    def validation_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        outputs, loss, bs = self._loss(batch)
        self.log("val_loss", loss, batch_size=bs, prog_bar=True, logger=True, on_step=False, on_epoch=True,
                 sync_dist=True, rank_zero_only=True)
        return self._predict(outputs.out)

    def test_step(self, batch: dict[str, Any], batch_idx: int) -> None:  # noqa: ARG002
        outputs, loss, bs = self._loss(batch)
        metrics: dict[str, Any] = {"test_loss": loss}
        _, iou = self._predict_and_score(outputs.out, batch["mask"].squeeze(1).long())
        metrics.update(iou)
        self.log_dict(metrics, batch_size=bs, prog_bar=False, logger=True, on_step=False, rank_zero_only=True)
