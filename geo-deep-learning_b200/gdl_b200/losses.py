"""Segmentation losses on the B200 kernels (csrc/loss_optim.cu), mirroring the callables the
reference's YAMLs instantiate: `segmentation_models_pytorch.losses.DiceLoss`,
`SoftCrossEntropyLoss` (configs/*.yaml, notebooks/00_quickstart.ipynb cell 15) and
`torch.nn.CrossEntropyLoss` (tests/test_notebooks_00quickstart.py:63).

Each module takes logits (N, K, H, W) and a target (N, H, W) / (N, 1, H, W) and returns a
scalar; the backward is a hand-written kernel wrapped in an autograd.Function.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .ops import LossSpec


def _nhwc_logits(logits: torch.Tensor) -> torch.Tensor:
    if logits.dim() != 4:
        raise ValueError(f"logits must be (N,K,H,W), got {tuple(logits.shape)}")
    x = logits.permute(0, 2, 3, 1)
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    return x


def _target(t: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    if t.dim() == 4 and t.shape[1] == 1:
        t = t[:, 0]
    if tuple(t.shape) != (n, h, w):
        raise ValueError(f"target shape {tuple(t.shape)} does not match logits ({n},{h},{w})")
    if t.dtype not in (torch.int64, torch.uint8):
        t = t.long()
    return t.contiguous()


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, target: torch.Tensor, spec: LossSpec) -> torch.Tensor:
        x = _nhwc_logits(logits)
        n, h, w, _ = x.shape
        t = _target(target, n, h, w)
        coeff, _ = ops.seg_loss_fwd(x, t, spec)
        ctx.save_for_backward(x, t, coeff)
        ctx.spec = spec
        return coeff[0].clone()

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        x, t, coeff = ctx.saved_tensors
        d = torch.empty_like(x)
        gs = grad_out.reshape(1).float().contiguous()
        ops.seg_loss_bwd(x, t, ctx.spec, coeff, gs, d)
        return d.permute(0, 3, 1, 2), None, None


class SegLoss(nn.Module):
    def __init__(self, spec: LossSpec) -> None:
        super().__init__()
        self.spec = spec

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _SegLossFn.apply(logits, target, self.spec)


class CrossEntropyLoss(SegLoss):
    """torch.nn.CrossEntropyLoss(reduction="mean", ignore_index, label_smoothing) for dense targets."""

    def __init__(self, ignore_index: int = -100, label_smoothing: float = 0.0) -> None:
        super().__init__(LossSpec(1.0, 0.0, label_smoothing, False, ignore_index))


class SoftCrossEntropyLoss(SegLoss):
    """smp.losses.SoftCrossEntropyLoss(reduction="mean", smooth_factor, ignore_index=-100)."""

    def __init__(self, reduction: str = "mean", smooth_factor: float | None = None, ignore_index: int | None = -100,
                 dim: int = 1) -> None:
        if reduction != "mean" or dim != 1:
            raise NotImplementedError("SoftCrossEntropyLoss: only reduction='mean', dim=1")
        super().__init__(LossSpec(1.0, 0.0, smooth_factor or 0.0, True, ignore_index))


class DiceLoss(SegLoss):
    """smp.losses.DiceLoss(mode, classes=None, log_loss=False, from_logits=True, smooth, ignore_index, eps)."""

    def __init__(self, mode: str, classes=None, log_loss: bool = False, from_logits: bool = True, smooth: float = 0.0,
                 ignore_index: int | None = None, eps: float = 1e-7) -> None:
        if mode not in ("binary", "multiclass"):
            raise NotImplementedError(f"DiceLoss mode {mode!r}")
        if classes is not None or log_loss or not from_logits:
            raise NotImplementedError("DiceLoss: classes / log_loss / from_logits=False are not implemented")
        self.mode = mode
        super().__init__(LossSpec(0.0, 1.0, 0.0, False, ignore_index, smooth, eps))

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.mode == "binary" and logits.shape[1] != 1:
            raise ValueError("binary DiceLoss expects a single logit channel")
        return super().forward(logits, target)
