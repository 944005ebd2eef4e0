"""Sliding-window raster inference (BASELINE configs[4]: 512-pixel tiles, stride 256, tile-sharded).

The reference ships no such driver (`tools/script_model.py` only exports a traced model); the per-tile computation is
exactly its eval path: `normalization` / `standardization` (utils/tensors.py:10-35) -> `model(x)` -> `softmax(dim=1)
.argmax(dim=1)` (segmentation_segformer.py:268-271).  Overlapping windows are blended by summing the fp32 logits of every
window that covers a pixel (argmax is invariant to the per-pixel window count; a single-logit model is blended with the
mean logit, so that `sigmoid > threshold` means the same for every threshold).

Everything on the device runs on the kernels of libgdlb200.so: uint8 HWC -> 16-bit NHWC normalisation, the model's eval
forward (`model.run` on an Engine with training=False), and the argmax kernel.  Host-side torch is used for the crop /
scatter-add glue only.  Multi-GPU: windows are dealt round-robin to the ranks, the partial logit sums are all-reduced once.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .engine import Act, Engine


def window_origins(size: int, tile: int, stride: int) -> list[int]:
    """Window starts along one axis: 0, stride, 2*stride, ... plus a last window flush with the border."""
    if size <= tile:
        return [0]
    xs = list(range(0, size - tile + 1, stride))
    if xs[-1] != size - tile:
        xs.append(size - tile)
    return xs


class SlidingWindowSegmenter:
    def __init__(self, model: torch.nn.Module, *, tile: int = 512, stride: int = 256, batch: int = 16, mean=None,
                 std=None, image_max: float = 255.0, threshold: float = 0.5, process_group=None,
                 cuda_graph: bool = False) -> None:
        if tile % 32:
            raise ValueError("tile must be divisible by 32")
        if not 0 < stride <= tile:
            raise ValueError("stride must be in (0, tile]")
        self.model, self.tile, self.stride, self.batch = model, tile, stride, batch
        self.image_max, self.threshold = image_max, threshold
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        dev = next(model.parameters()).device
        self.dev = dev
        self.mean = torch.as_tensor(mean, dtype=torch.float32, device=dev) if mean is not None else None
        self.std = torch.as_tensor(std, dtype=torch.float32, device=dev) if std is not None else None
        self.windows_done = 0
        # A window batch is ~1300 launches of 5-50 us kernels (SegFormer-B5): issued from Python they are launch-bound.
        # cuda_graph=True captures normalise + forward for full batches once (after one eager batch) and replays it; the
        # ragged last batch runs eagerly.  Same kernels either way.
        self.cuda_graph = cuda_graph
        self._graph = None
        self._static_in: torch.Tensor | None = None
        self._static_out: torch.Tensor | None = None
        self._eager_batches = 0

    def _forward_eager(self, crops: torch.Tensor) -> torch.Tensor:
        model = self.model
        c = crops.shape[3]
        x16 = ops.normalize_to_nhwc(crops, False, model.compute_dtype, (c + 7) // 8 * 8, self.mean, self.std, self.image_max)
        eng = Engine(model.compute_dtype, training=False, wcache=model._wcache)
        x = Act(x16, needs_grad=False)
        return model.run(eng, x, c) if getattr(model, "needs_bands", False) else model.run(eng, x)   # fp32 (B, t, t, K)

    def _forward(self, crops: torch.Tensor) -> torch.Tensor:
        """(B, t, t, C) uint8 window batch -> fp32 logits (B, t, t, K); the result is only valid until the next call"""
        if not (self.cuda_graph and crops.is_cuda and crops.shape[0] == self.batch):
            return self._forward_eager(crops)
        if self._static_in is None or self._static_in.shape != crops.shape:
            self._static_in, self._graph, self._eager_batches = torch.empty_like(crops), None, 0
        self._static_in.copy_(crops)
        if self._graph is None:
            if self._eager_batches == 0:  # lazy CUDA init, function attributes, packed weights: not inside a capture
                self._eager_batches = 1
                return self._forward_eager(self._static_in)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._static_out = self._forward_eager(self._static_in)
            self._graph = graph
        self._graph.replay()
        return self._static_out

    @torch.no_grad()
    def logits(self, raster: torch.Tensor) -> torch.Tensor:
        """raster: (H, W, C) uint8 (device, or host — it is copied once).  Returns the summed fp32 logits (H, W, K)."""
        if raster.dim() != 3 or raster.dtype != torch.uint8:
            raise ValueError("raster must be a (H, W, C) uint8 tensor")
        model, t = self.model, self.tile
        raster = raster.to(self.dev, non_blocking=True)
        h, w, c = raster.shape
        hp, wp = max(h, t), max(w, t)
        if (hp, wp) != (h, w):  # a raster smaller than one window is zero padded (bottom / right)
            padded = torch.zeros((hp, wp, c), dtype=torch.uint8, device=self.dev)
            padded[:h, :w] = raster
            raster = padded
        wins = [(y, x) for y in window_origins(hp, t, self.stride) for x in window_origins(wp, t, self.stride)]
        mine = wins[self.rank::self.world]
        acc = None
        for i in range(0, len(mine), self.batch):
            chunk = mine[i:i + self.batch]
            crops = torch.stack([raster[y:y + t, x:x + t] for y, x in chunk])            # (B, t, t, C) uint8
            out = self._forward(crops)                                                    # fp32 (B, t, t, K)
            if acc is None:
                acc = torch.zeros((hp, wp, out.shape[3]), dtype=torch.float32, device=self.dev)
            for j, (y, x) in enumerate(chunk):
                acc[y:y + t, x:x + t] += out[j]
            self.windows_done += len(chunk)
        if acc is None:  # more ranks than windows: this rank contributes zeros
            k = getattr(model, "num_classes", None) or getattr(model, "classes", None)
            if k is None:
                raise RuntimeError("no window on this rank and the model does not expose its class count")
            acc = torch.zeros((hp, wp, k), dtype=torch.float32, device=self.dev)
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
        if acc.shape[2] == 1:
            # one logit: sigmoid(sum) > threshold equals sigmoid(mean) > threshold only at threshold 0.5, so blend with
            # the MEAN logit (divide by the number of windows covering the pixel); argmax (K > 1) is invariant to it
            cnt = torch.zeros((hp, wp, 1), dtype=torch.float32, device=self.dev)
            for y, x in wins:
                cnt[y:y + t, x:x + t] += 1.0
            acc = acc / cnt.clamp_(min=1.0)
        return acc[:h, :w]

    @torch.no_grad()
    def predict(self, raster: torch.Tensor) -> torch.Tensor:
        """(H, W, C) uint8 raster -> (H, W) uint8 class map (multi-class: argmax; one logit: sigmoid > threshold)."""
        acc = self.logits(raster)
        h, w, k = acc.shape
        cls = ops.argmax_classes(acc.contiguous().view(1, h, w, k), self.threshold)
        return cls.view(h, w).to(torch.uint8)
