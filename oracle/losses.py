"""ORACLE (test infrastructure). Loss restatements.

`segmentation_models_pytorch==0.5.0` is not in /root/reference (un-vendored, pinned at
uv.lock.cu128:3250), so DiceLoss / SoftCrossEntropyLoss are restated from the published smp
0.5.0 semantics (losses/dice.py, losses/soft_ce.py, losses/_functional.py):

  dice:    p = log_softmax(x,1).exp() (multiclass) | logsigmoid(x).exp() (binary); one-hot target;
           sums over (batch, pixels); score = (2*I + smooth) / (C + smooth).clamp_min(eps);
           loss = (1 - score) * [sum(t) > 0]; mean over classes.
  soft-ce: (1-eps) * nll + eps/K * (-sum_c log p_c), mean over ALL pixels (ignored -> 0).

PARITY STATUS: unpinned for values (the reference has no test or golden vector for any loss);
CrossEntropyLoss is torch's own and needs no restatement.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dice_loss(logits: torch.Tensor, target: torch.Tensor, mode: str, smooth: float = 0.0, eps: float = 1e-7,
              ignore_index: int | None = None) -> torch.Tensor:
    bs = target.size(0)
    k = logits.size(1)
    dims = (0, 2)
    if mode == "multiclass":
        p = F.log_softmax(logits, dim=1).exp()
        t = target.view(bs, -1).long()
        p = p.view(bs, k, -1)
        if ignore_index is not None:
            mask = t != ignore_index
            p = p * mask.unsqueeze(1)
            t1 = F.one_hot((t * mask).long(), k).permute(0, 2, 1) * mask.unsqueeze(1)
        else:
            t1 = F.one_hot(t, k).permute(0, 2, 1)
    elif mode == "binary":
        p = F.logsigmoid(logits).exp().view(bs, 1, -1)
        t1 = target.view(bs, 1, -1)
        if ignore_index is not None:
            mask = t1 != ignore_index
            p = p * mask
            t1 = t1 * mask
    else:
        raise ValueError(mode)
    t1 = t1.type_as(p)
    inter = torch.sum(p * t1, dim=dims)
    card = torch.sum(p + t1, dim=dims)
    score = (2.0 * inter + smooth) / (card + smooth).clamp_min(eps)
    loss = 1.0 - score
    loss = loss * (t1.sum(dims) > 0).to(loss.dtype)
    return loss.mean()


def soft_ce_loss(logits: torch.Tensor, target: torch.Tensor, smooth_factor: float, ignore_index: int | None = -100):
    lp = F.log_softmax(logits, dim=1)
    t = target.long().unsqueeze(1)
    if ignore_index is not None:
        pad = t.eq(ignore_index)
        t = t.masked_fill(pad, 0)
        nll = -lp.gather(1, t).masked_fill(pad, 0.0)
        smooth = -lp.sum(dim=1, keepdim=True).masked_fill(pad, 0.0)
    else:
        nll = -lp.gather(1, t)
        smooth = -lp.sum(dim=1, keepdim=True)
    nll, smooth = nll.squeeze(1).mean(), smooth.squeeze(1).mean()
    e = smooth_factor / lp.size(1)
    return (1.0 - smooth_factor) * nll + e * smooth
