"""Device-side replacements of the two host-side pieces the reference tasks wrap around the model:

* `on_before_batch_transfer` runs a kornia pipeline on the CPU batch (segmentation_segformer.py:95-125,206-216 and the
  UNet++ / DOFA twins).  Here the same augmentation is `on_after_batch_transfer`: the draws are made on the host, the
  pixels are moved by `gdl_augment_normalize` on the device (gdl_b200/augment.py).
* `test_step` feeds the argmax map to torchmetrics' `MeanIoU` (segmentation_segformer.py:283-296).  Here the confusion
  counts come out of the argmax kernel (`gdl_argmax_confusion`, gdl_b200/metrics.py).
"""
from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from .. import ops
from ..augment import BatchAugmenter
from ..metrics import MeanIoU


class GpuSideHooks:
    gpu_augment: bool = True  # set False to keep a host-side augmentation of the caller's own

    def _is_training(self) -> bool:
        try:
            return bool(self.trainer.training)  # what the reference checks (segmentation_segformer.py:212)
        except Exception:  # noqa: BLE001 - not attached to a Lightning Trainer
            return bool(self.training)

    def on_after_batch_transfer(self, batch: dict[str, Any], dataloader_idx: int) -> dict[str, Any]:  # noqa: ARG002
        if not (self.gpu_augment and self._is_training()):
            return batch
        image, mask = batch["image"], batch["mask"]
        ops.require_cuda(image, "the batch augmentation (on_after_batch_transfer)")
        aug = getattr(self, "_augmenter", None)
        if aug is None or (aug.h, aug.w) != tuple(image.shape[2:4]):
            aug = self._augmenter = BatchAugmenter(tuple(image.shape[2:4]))
        m3 = mask.reshape(mask.shape[0], *mask.shape[-2:])
        if m3.dtype not in (torch.int64, torch.uint8):
            m3 = m3.long()
        new_image, new_mask = aug(image.float().contiguous(), m3.contiguous(), chw=True, out_dtype=torch.float32)
        batch.update({"image": new_image, "mask": new_mask.reshape(mask.shape).to(mask.dtype)})
        return batch

    def _iou_metric(self) -> MeanIoU:
        m = getattr(self, "_mean_iou", None)
        if m is None:
            m = self._mean_iou = MeanIoU(len(self.labels), per_class=True, include_background=True, labels=self.labels)
        return m

    def _predict_and_score(self, y_hat: Tensor, target: Tensor) -> tuple[Tensor, dict[str, Tensor]]:
        """logits (N,K,H,W) + target (N,H,W) -> (class map, {"meaniou_<label>": IoU}) of this batch, as
        `self.iou_classwise_metric(y_hat, y)` followed by `.reset()` in the reference's test_step."""
        nhwc = y_hat.permute(0, 2, 3, 1)
        if nhwc.dtype != torch.float32 or not nhwc.is_contiguous():
            nhwc = nhwc.float().contiguous()
        metric = self._iou_metric()
        classes = metric.update(nhwc, target.contiguous(), self.threshold)
        values = metric.compute()
        metric.reset()
        return classes, values
