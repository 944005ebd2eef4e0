"""GPU-side batch augmentation — the step right in front of the hot path.

The reference builds a kornia pipeline in every task's `_apply_aug` and runs it on the HOST, on the float batch, in
`on_before_batch_transfer` (segmentation_segformer.py:95-125,206-216; segmentation_unetplus.py:92-122;
segmentation_dofa.py:91-121):

    AugmentationSequential(RandomHorizontalFlip(p=.5), RandomVerticalFlip(p=.5),
                           RandomRotation90(times=(1, 3), p=.5, align_corners=True),
                           RandomResizedCrop(size, scale=(1, 2), p=.5, align_corners=False),      # "zoom in"
                           RandomResizedCrop(size, scale=(.5, 1), p=.5, align_corners=False),     # "zoom out"
                           data_keys=None, random_apply=1)

Here only the random DRAWS stay on the host (a few integers per sample); the pixels are moved by one kernel
(`gdl_augment_normalize`) that is fused with the uint8 -> 16-bit NHWC patch normalisation, so augmentation costs no extra
pass over the batch and no host float work.

kornia (pinned >=0.8,<0.9, `pyproject.toml`) is not vendored in the reference tree and not installable offline, so the
sampling rules below restate its published behaviour (PARITY UNPINNED for the random streams; the pixel arithmetic is
pinned to torch's flip / rot90 / F.interpolate in tests):
  * `random_apply=1`: ONE of the five operations is drawn uniformly per batch; it is applied to each sample
    independently with its own probability p.
  * RandomRotation90(times=(1, 3)): an integer number of quarter turns in [1, 3] per sample (kornia warps by 90*times
    degrees about the centre with align_corners=True, which is an exact pixel permutation: torch.rot90).
  * RandomResizedCrop: up to 10 tries of area = U(scale) * H * W, log-ratio = U(log 3/4, log 4/3),
    w = floor(round(sqrt(area * ratio))), h = floor(round(sqrt(area / ratio))), accepted when 0 < w < W and 0 < h < H;
    otherwise the whole tile (the tile's ratio 1 lies inside [3/4, 4/3]).  With scale=(1, 2) no try can be accepted,
    so that entry is an identity.  The window origin is floor(U(0, W - w + 1)), floor(U(0, H - h + 1)); the window is
    resized back to `size`: image bilinear (align_corners=False), mask nearest.
"""
from __future__ import annotations

import math

import torch

from . import ops

OPS = ("hflip", "vflip", "rot90", "resized_crop_zoom_in", "resized_crop_zoom_out")


class BatchAugmenter:
    """Draws the per-sample augmentation parameters on the host and applies them on the device.

    `sample(n)` returns the int32 (n, 6) parameter table {op, k, y0, x0, ch, cw} of `gdl_augment_normalize`;
    `__call__` runs the fused augment + normalise kernel."""

    def __init__(self, image_size: tuple[int, int], *, p: float = 0.5, times: tuple[int, int] = (1, 3),
                 zoom_in_scale: tuple[float, float] = (1.0, 2.0), zoom_out_scale: tuple[float, float] = (0.5, 1.0),
                 ratio: tuple[float, float] = (3.0 / 4.0, 4.0 / 3.0), generator: torch.Generator | None = None) -> None:
        self.h, self.w = int(image_size[0]), int(image_size[1])
        if self.h <= 0 or self.w <= 0:
            raise ValueError("image_size must be positive")
        if not 0.0 <= p <= 1.0:
            raise ValueError("p must be in [0, 1]")
        self.p, self.times = p, times
        self.scales = {"resized_crop_zoom_in": zoom_in_scale, "resized_crop_zoom_out": zoom_out_scale}
        self.ratio = ratio
        self.gen = generator
        self.last_op: str | None = None

    # -- host-side random draws ---------------------------------------------------------------
    def _rand(self, *shape: int) -> torch.Tensor:
        return torch.rand(*shape, generator=self.gen)

    def _crop_boxes(self, n: int, scale: tuple[float, float]) -> torch.Tensor:
        """(n, 4) int64 {y0, x0, ch, cw} following kornia's ResizedCropGenerator (10 tries, whole-tile fallback)."""
        h, w = self.h, self.w
        area = (self._rand(n, 10) * (scale[1] - scale[0]) + scale[0]) * (h * w)
        lo, hi = math.log(self.ratio[0]), math.log(self.ratio[1])
        aspect = torch.exp(self._rand(n, 10) * (hi - lo) + lo)
        cw = torch.sqrt(area * aspect).round().floor()
        ch = torch.sqrt(area / aspect).round().floor()
        ok = (cw > 0) & (cw < w) & (ch > 0) & (ch < h)
        first = ok.float().argmax(1)                       # index of the first accepted try (0 if none)
        rows = torch.arange(n)
        any_ok = ok.any(1)
        cw = torch.where(any_ok, cw[rows, first], torch.full((n,), float(w)))
        ch = torch.where(any_ok, ch[rows, first], torch.full((n,), float(h)))
        x0 = torch.floor(self._rand(n) * (w - cw + 1)).clamp_(min=0)
        y0 = torch.floor(self._rand(n) * (h - ch + 1)).clamp_(min=0)
        x0 = torch.minimum(x0, w - cw)
        y0 = torch.minimum(y0, h - ch)
        return torch.stack([y0, x0, ch, cw], 1).long()

    def sample(self, n: int) -> torch.Tensor:
        """int32 (n, 6) host tensor; also records the batch's operation in `last_op`."""
        op_idx = int(torch.randint(0, len(OPS), (1,), generator=self.gen))
        name = OPS[op_idx]
        self.last_op = name
        params = torch.zeros((n, 6), dtype=torch.int32)
        apply = self._rand(n) < self.p
        if name == "hflip":
            params[:, 0] = torch.where(apply, ops.AUG_HFLIP, ops.AUG_IDENTITY)
        elif name == "vflip":
            params[:, 0] = torch.where(apply, ops.AUG_VFLIP, ops.AUG_IDENTITY)
        elif name == "rot90":
            if self.h != self.w:
                raise ValueError("RandomRotation90 needs square tiles")
            k = torch.randint(self.times[0], self.times[1] + 1, (n,), generator=self.gen)
            params[:, 0] = torch.where(apply, ops.AUG_ROT90, ops.AUG_IDENTITY)
            params[:, 1] = torch.where(apply, k, 0).int()
        else:
            box = self._crop_boxes(n, self.scales[name])
            whole = (box[:, 2] == self.h) & (box[:, 3] == self.w)
            crop = apply & ~whole                          # a whole-tile window resized to itself is the identity
            params[:, 0] = torch.where(crop, ops.AUG_CROP, ops.AUG_IDENTITY)
            params[:, 2:6] = torch.where(crop[:, None], box, torch.zeros_like(box)).int()
        return params

    # -- device application -------------------------------------------------------------------
    def __call__(self, image: torch.Tensor, mask: torch.Tensor | None, *, chw: bool, out_dtype: torch.dtype,
                 ld: int = 0, mean: torch.Tensor | None = None, std: torch.Tensor | None = None,
                 image_max: float = 0.0, params: torch.Tensor | None = None):
        """image: uint8 / f32 batch on the device (NHWC if chw=False else NCHW); mask (N,H,W) int64/uint8 or None.
        Returns (image', mask') as `ops.augment_normalize`."""
        n = image.shape[0]
        hw = tuple(image.shape[2:4]) if chw else tuple(image.shape[1:3])
        if hw != (self.h, self.w):
            raise ValueError(f"batch tiles are {hw}, augmenter was built for {(self.h, self.w)}")
        if params is None:
            params = self.sample(n)
        if not params.is_cuda:
            params = params.pin_memory().to(image.device, non_blocking=True) if torch.cuda.is_available() \
                else params.to(image.device)
        return ops.augment_normalize(image, chw, mask, params, out_dtype, ld, mean, std, image_max)
