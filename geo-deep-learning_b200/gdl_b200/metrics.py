"""Evaluation metric of the tasks' test_step on the device.

The reference computes `torchmetrics.segmentation.MeanIoU(num_classes, per_class=True, input_format="index",
include_background=True)` wrapped in `ClasswiseWrapper(labels)` from the argmax map (segmentation_segformer.py:78-92,
283-296; same in the UNet++ and DOFA tasks).  torchmetrics one-hot encodes both maps and reduces per sample; here the
per-sample confusion counts come out of the same kernel that takes the argmax (`gdl_argmax_confusion`), and the metric is
a few (N, K) integer reductions of them.  torchmetrics 1.8 semantics (restated; the package is not vendored):
per sample and class, IoU = intersection / union; classes with an empty union in a sample are skipped for that sample;
`compute()` = mean over the samples in which the class was present.
"""
from __future__ import annotations

import torch

from . import ops


class MeanIoU:
    """Accumulating per-class IoU with the update/compute/reset surface of the torchmetrics object it replaces."""

    def __init__(self, num_classes: int, per_class: bool = True, include_background: bool = True,
                 labels: list[str] | None = None, prefix: str = "meaniou_") -> None:
        if num_classes < 2:
            raise ValueError("num_classes must be >= 2 (the tasks use 2 for a single-logit model)")
        self.num_classes, self.per_class, self.include_background = num_classes, per_class, include_background
        self.labels = labels if labels is not None else [str(i) for i in range(num_classes)]
        if len(self.labels) != num_classes:
            raise ValueError("one label per class expected")
        self.prefix = prefix
        self.score: torch.Tensor | None = None
        self.count: torch.Tensor | None = None

    def reset(self) -> None:
        self.score = self.count = None

    def update_from_confusion(self, conf: torch.Tensor) -> None:
        """conf: (N, K, K) int64, conf[n][target][prediction]."""
        if conf.dim() != 3 or conf.shape[1] != self.num_classes or conf.shape[2] != self.num_classes:
            raise ValueError(f"confusion counts must be (N, {self.num_classes}, {self.num_classes})")
        inter = conf.diagonal(dim1=1, dim2=2).double()
        union = conf.sum(1).double() + conf.sum(2).double() - inter
        if not self.include_background:
            inter, union = inter[:, 1:], union[:, 1:]
        valid = union > 0
        score = torch.where(valid, inter / union.clamp_min(1), torch.zeros_like(inter))
        s, c = score.sum(0), valid.sum(0)
        self.score = s if self.score is None else self.score + s
        self.count = c if self.count is None else self.count + c

    def update(self, logits_nhwc: torch.Tensor, target: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
        """logits fp32 (N,H,W,K) on the device, target (N,H,W) int64/uint8.  Returns the class map (N,H,W) int64."""
        classes, conf = ops.argmax_confusion(logits_nhwc, target, threshold)
        self.update_from_confusion(conf)
        return classes

    def compute(self):
        if self.score is None:
            raise RuntimeError("MeanIoU.compute() called before update()")
        per_class = torch.nan_to_num(self.score / self.count, nan=0.0).float()
        if not self.per_class:
            return per_class.mean()
        labels = self.labels if self.include_background else self.labels[1:]
        return {f"{self.prefix}{name}": per_class[i] for i, name in enumerate(labels)}  # ClasswiseWrapper naming

    def __call__(self, logits_nhwc: torch.Tensor, target: torch.Tensor, threshold: float = 0.5):
        """forward() of a torchmetrics Metric: the value of THIS batch, while the running state also accumulates it."""
        classes, conf = ops.argmax_confusion(logits_nhwc, target, threshold)
        batch = MeanIoU(self.num_classes, self.per_class, self.include_background, self.labels, self.prefix)
        batch.update_from_confusion(conf)
        self.update_from_confusion(conf)
        return batch.compute()
