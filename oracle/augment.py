"""ORACLE (test infrastructure). The pixel arithmetic of the reference's host-side kornia augmentation
(`_apply_aug`, segmentation_segformer.py:95-125 and its twins) restated with torch's own operators on the float batch:

  RandomHorizontalFlip / RandomVerticalFlip -> torch.flip over W / H
  RandomRotation90(times)                   -> torch.rot90(x, times, dims=(H, W))  (kornia warps by 90*times degrees about
                                               the centre with align_corners=True: an exact pixel permutation)
  RandomResizedCrop(size, align_corners=False, cropping_mode="slice")
                                            -> x[..., y0:y0+ch, x0:x0+cw] resized with F.interpolate: image bilinear
                                               (align_corners=False), mask nearest

PARITY STATUS: kornia (>=0.8,<0.9, un-vendored, not installable offline) cannot be executed here, so agreement with kornia
itself is UNPINNED; what is pinned is that the CUDA kernel equals these torch operators given the same draws.  The
rotation direction of kornia's RandomRotation90 is not recoverable from the reference tree; both directions are covered by
times in {1, 2, 3}, so the distribution of results is the same.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IDENTITY, HFLIP, VFLIP, ROT90, CROP = range(5)


def apply_params(image: torch.Tensor, mask: torch.Tensor | None, params: torch.Tensor):
    """image (N,C,H,W) float, mask (N,H,W) integer or None, params (N,6) int {op,k,y0,x0,ch,cw} -> (image', mask')."""
    n, _, h, w = image.shape
    out_i = torch.empty_like(image)
    out_m = torch.empty_like(mask) if mask is not None else None
    for s in range(n):
        op, k, y0, x0, ch, cw = (int(v) for v in params[s])
        img = image[s]
        msk = mask[s] if mask is not None else None
        if op == HFLIP:
            img = img.flip(-1)
            msk = msk.flip(-1) if msk is not None else None
        elif op == VFLIP:
            img = img.flip(-2)
            msk = msk.flip(-2) if msk is not None else None
        elif op == ROT90:
            img = torch.rot90(img, k, dims=(-2, -1))
            msk = torch.rot90(msk, k, dims=(-2, -1)) if msk is not None else None
        elif op == CROP:
            img = F.interpolate(img[None, :, y0:y0 + ch, x0:x0 + cw], size=(h, w), mode="bilinear",
                                align_corners=False)[0]
            if msk is not None:
                m = msk[None, None, y0:y0 + ch, x0:x0 + cw]
                msk = F.interpolate(m.double(), size=(h, w), mode="nearest")[0, 0].to(mask.dtype)
        out_i[s] = img
        if out_m is not None:
            out_m[s] = msk
    return out_i, out_m
