"""`SegmentationSegformer` on the B200 kernels — constructor keywords, hooks and batch contract of
geo_deep_learning/tasks_with_models/segmentation_segformer.py:32-295; `configure_model` builds
gdl_b200.models.segformer.SegFormer (same state_dict as SegFormerSegmentationModel), the training
step casts the mask to (B,H,W) int64 exactly like the reference (:224-225), eval post-processing
(`softmax(dim=1).argmax(dim=1)` / `sigmoid > 0.5`, :268-271) runs on the argmax kernel."""
from __future__ import annotations

from typing import Any, Callable

import torch
from torch import Tensor

from .. import ops
from ..models.segformer import SegFormer
from . import _common
from ._hooks import GpuSideHooks
from .segmentation_unetplus import _Base, _strip_model_prefix


class SegmentationSegformer(GpuSideHooks, _Base):
    def __init__(self, encoder: str, *, image_size: tuple[int, int], in_channels: int, num_classes: int,
                 max_samples: int, loss: Callable, optimizer: Callable = torch.optim.Adam,
                 scheduler: Callable | None = None, scheduler_config: dict[str, Any] | None = None,
                 use_dynamic_encoder: bool = False, freeze_layers: list[str] | None = None, weights: str | None = None,
                 class_labels: list[str] | None = None, class_colors: list[str] | None = None,
                 weights_from_checkpoint_path: str | None = None, compute_dtype: torch.dtype = torch.bfloat16,
                 **kwargs: object) -> None:
        super().__init__()
        self.save_hyperparameters(ignore=["loss"])
        self.encoder, self.in_channels, self.num_classes = encoder, in_channels, num_classes
        self.image_size, self.max_samples = image_size, max_samples
        self.loss, self.optimizer, self.scheduler = loss, optimizer, scheduler
        self.scheduler_config = scheduler_config or {"interval": "epoch"}
        self.use_dynamic_encoder, self.freeze_layers, self.weights = use_dynamic_encoder, freeze_layers, weights
        self.weights_from_checkpoint_path = weights_from_checkpoint_path
        self.class_colors = class_colors
        self.threshold = 0.5
        k = 2 if num_classes == 1 else num_classes
        self.labels = [str(i) for i in range(k)] if class_labels is None else class_labels
        self.compute_dtype = compute_dtype
        self.model: SegFormer | None = None

    def configure_model(self) -> None:
        if self.model is not None:
            return
        # `weights` reaches the model exactly as in the reference (segmentation_segformer.py:129-137): a pretrained request
        # is refused loudly there instead of being dropped here
        _common.check_pretrained_request(self.weights, "SegmentationSegformer")
        # the reference's train-mode regularisation: DropPath 0.1 in every MiT variant (mix_transformer.py:614-705) and
        # Dropout2d(0.1) in the MLP decoder (segformer_mlp.py:32,73)
        self.model = SegFormer(self.encoder, self.in_channels, self.weights, self.freeze_layers, self.num_classes,
                               use_dynamic_encoder=self.use_dynamic_encoder, compute_dtype=self.compute_dtype,
                               drop_path_rate=0.1, dropout_ratio=0.1)
        if self.weights_from_checkpoint_path:
            _common.load_weights_from_checkpoint(self.model, self.weights_from_checkpoint_path,
                                                 _common.hparam(self, "load_parts"),
                                                 trust_pickle=bool(_common.hparam(self, "trust_checkpoint_pickle", False)))

    def configure_optimizers(self):
        """segmentation_segformer.py:150-199 (OneCycleLR horizon from the trainer; tasks/_common.py)"""
        return _common.configure_optimizers(self)

    def forward(self, image: Tensor) -> Tensor:
        return self.model(image)

    def training_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        x, y = batch["image"], batch["mask"].squeeze(1).long()
        loss = self.loss(self(x), y)
        self.log("train_loss", loss, batch_size=x.shape[0], prog_bar=True, logger=True, on_step=False,
                 on_epoch=True, sync_dist=True, rank_zero_only=True)
        return loss

    def _predict(self, y_hat: Tensor) -> Tensor:
        nhwc = y_hat.permute(0, 2, 3, 1)
        if nhwc.dtype != torch.float32 or not nhwc.is_contiguous():
            nhwc = nhwc.float().contiguous()
        return ops.argmax_classes(nhwc, self.threshold)

    def validation_step(self, batch: dict[str, Any], batch_idx: int) -> Tensor:  # noqa: ARG002
        x, y = batch["image"], batch["mask"].squeeze(1).long()
        y_hat = self(x)
        self.log("val_loss", self.loss(y_hat, y), batch_size=x.shape[0], prog_bar=True, logger=True,
                 on_step=False, on_epoch=True, sync_dist=True, rank_zero_only=True)
        return self._predict(y_hat)

    def test_step(self, batch: dict[str, Any], batch_idx: int) -> None:  # noqa: ARG002
        x, y = batch["image"], batch["mask"].squeeze(1).long()
        y_hat = self(x)
        metrics: dict[str, Any] = {"test_loss": self.loss(y_hat, y)}
        _, iou = self._predict_and_score(y_hat, y)  # MeanIoU per class on the argmax kernel (:283-287)
        metrics.update(iou)
        self.log_dict(metrics, batch_size=x.shape[0], prog_bar=False, logger=True, on_step=False,
                      rank_zero_only=True)
