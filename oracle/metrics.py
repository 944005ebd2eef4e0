"""ORACLE (test infrastructure). The evaluation metric of the reference's test_step
(segmentation_segformer.py:78-92,283-296 and twins): torchmetrics `MeanIoU(num_classes, per_class=True,
input_format="index", include_background=True)` wrapped in `ClasswiseWrapper`.

torchmetrics (pinned >=1.8,<1.9) is un-vendored and not installable offline; this restates its published 1.8 algorithm
(functional/segmentation/mean_iou.py, segmentation/mean_iou.py) — PARITY UNPINNED against torchmetrics itself:
  update:  one-hot both index maps; per SAMPLE and class: intersection = sum(pred & target),
           union = sum(pred) + sum(target) - intersection; score = intersection / union (0 where union == 0);
           valid = union > 0; state.score += sum_n(score * valid); state.num_batches += sum_n(valid)
  compute: score / num_batches per class (nan -> 0 for a class never seen).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def confusion_per_sample(pred: torch.Tensor, target: torch.Tensor, num_classes: int) -> torch.Tensor:
    """(N,H,W) int maps -> (N,K,K) int64 counts conf[n][target][prediction] (plain bincount)."""
    n = pred.shape[0]
    idx = target.reshape(n, -1).long() * num_classes + pred.reshape(n, -1).long()
    return torch.stack([torch.bincount(idx[i], minlength=num_classes * num_classes) for i in range(n)]).view(
        n, num_classes, num_classes)


def mean_iou_update(pred: torch.Tensor, target: torch.Tensor, num_classes: int):
    """-> (score_sum (K,), valid_count (K,)) of one batch, exactly as torchmetrics accumulates them."""
    p = F.one_hot(pred.long(), num_classes).movedim(-1, 1)
    t = F.one_hot(target.long(), num_classes).movedim(-1, 1)
    dims = list(range(2, p.dim()))
    inter = (p & t).sum(dim=dims).double()
    union = p.sum(dim=dims).double() + t.sum(dim=dims).double() - inter
    score = torch.where(union > 0, inter / union.clamp_min(1), torch.zeros_like(inter))
    valid = union > 0
    return (score * valid).sum(0), valid.sum(0)


def mean_iou_compute(score_sum: torch.Tensor, valid_count: torch.Tensor) -> torch.Tensor:
    return torch.nan_to_num(score_sum / valid_count, nan=0.0)
