"""Fused training step: uint8 tiles in, updated parameters out, no autograd in between.

normalise (uint8 HWC -> 16-bit NHWC) -> engine forward -> loss kernels -> engine backward ->
(NCCL all-reduce of the flat gradient) -> (clip) -> fused Adam on flat fp32 buffers.
This is the path bench.py times; the autograd.Function route in models/ is the Lightning
drop-in and runs the same kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .engine import Act, Engine, refresh_packed_weights
from .ops import LossSpec


def gradient_buckets(numels: list[int], keys: list, nb: int = 3) -> list[tuple[int, int, frozenset]]:
    """Contiguous [start, end) ranges of the flat gradient buffer (parameters in `numels` order, each slot rounded up to 4
    elements as FusedTrainer lays them out), about `nb` of equal size, returned in BACKWARD order: the tail of the buffer —
    the head / decoder, whose gradients are complete first — comes first.  Every element and every key belongs to exactly
    one bucket."""
    offsets, total = [], 0
    for n in numels:
        offsets.append(total)
        total += (n + 3) // 4 * 4
    bounds = [total * (i + 1) // nb for i in range(nb)]
    buckets, start, ids, bi = [], 0, [], 0
    for n, off, key in zip(numels, offsets, keys):
        ids.append(key)
        end = off + (n + 3) // 4 * 4
        if bi < nb - 1 and end >= bounds[bi]:
            buckets.append((start, end, frozenset(ids)))
            start, ids = end, []
            while bi < nb - 1 and end >= bounds[bi]:
                bi += 1
    if ids:
        buckets.append((start, total, frozenset(ids)))
    buckets.reverse()
    return buckets


class FusedTrainer:
    def __init__(self, model: torch.nn.Module, loss: LossSpec, *, lr: float = 1e-4, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, mean=None, std=None, image_max: float = 255.0,
                 clip_grad_norm: float | None = None, process_group=None, sync_bn: bool = False,
                 acc_dtype: torch.dtype = torch.float32, cuda_graph: bool = False, input_chw: bool = False,
                 loss_scale: float | str | None = None, growth_interval: int = 2000) -> None:
        self.model = model
        self.loss = loss
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.image_max = image_max
        self.input_chw = input_chw  # tiles arrive (N,C,H,W) — the layout of the WebDataset shards — instead of (N,H,W,C)
        # fp16 loss scaling — what Lightning's `precision: 16-mixed` GradScaler does around the reference's backward
        # (SURVEY §8 a20): d(loss) is multiplied by S before the backward, the flat gradient by 1/S before clip / Adam.
        # None = off (bf16 needs none); a float = static S; "dynamic" = torch.amp.GradScaler's rule (S0 = 65536, a step with
        # a non-finite gradient is skipped and halves S, `growth_interval` clean steps double it).
        if isinstance(loss_scale, str) and loss_scale != "dynamic":
            raise ValueError("loss_scale must be None, a number or 'dynamic'")
        if loss_scale == "dynamic" and cuda_graph:
            raise NotImplementedError("dynamic loss scaling skips steps on the host: use a static scale with cuda_graph=True")
        self.loss_scale_mode = loss_scale
        self.growth_interval, self._good_steps, self.skipped_steps = growth_interval, 0, 0
        self.clip = clip_grad_norm
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        if self.world > 1 and self.group is None:
            self.group = dist.group.WORLD
        self.sync_bn = sync_bn and self.world > 1
        # SyncBN statistics over NVLink peer memory (one small kernel per layer) where torch's symmetric memory is available;
        # otherwise NCCL all-reduces.  Every rank takes the same decision (the outcome is agreed on with an all-reduce).
        self.bn_exchange = None
        self.bn_exchange_kind = "nccl" if self.sync_bn else None
        self.acc_dtype = acc_dtype  # fp32 on the GPU; float64 only in CPU host-logic tests
        self.step_count = 0
        dev = next(model.parameters()).device
        self.dev = dev
        self.params = [p for p in model.parameters() if p.requires_grad]
        # every parameter starts on a 16-byte boundary of the flat buffers: the wgrad kernel accumulates pointwise-conv
        # gradients straight into them with 128-bit reductions (the gaps stay zero: Adam leaves them at zero)
        offsets, total = [], 0
        for p in self.params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=acc_dtype, device=dev)
        self.gflat = torch.zeros(total, dtype=acc_dtype, device=dev)
        self.m = torch.zeros(total, dtype=acc_dtype, device=dev)
        self.v = torch.zeros(total, dtype=acc_dtype, device=dev)
        self.grad_dst: dict[int, torch.Tensor] = {}
        with torch.no_grad():
            for p, off in zip(self.params, offsets):
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                g = self.gflat[off:off + n].view(p.shape)
                p.grad = g
                self.grad_dst[id(p)] = g
        # gradient buckets for the overlapped all-reduce (DDP's bucketed reduction, configs/segformer_config_RGB.yaml:6-14):
        # contiguous ranges of the flat buffer in BACKWARD order (the tail of the flat buffer = the layers whose gradients are
        # complete first).  A bucket is all-reduced on NCCL's stream as soon as every parameter in it has its gradient, while
        # the backward of the earlier layers keeps the GPU busy; what is left is reduced in optimizer_step.
        self.overlap_allreduce = self.world > 1 and ops.option("overlap_allreduce") != 0
        self._buckets = gradient_buckets([p.numel() for p in self.params], [id(p) for p in self.params]) \
            if self.overlap_allreduce else []
        self._pending: list = []
        self._next_bucket = 0
        self.mean = torch.as_tensor(mean, dtype=acc_dtype, device=dev) if mean is not None else None
        self.std = torch.as_tensor(std, dtype=acc_dtype, device=dev) if std is not None else None
        self.loss_scale = (None if loss_scale is None else
                           torch.full((1,), 65536.0 if loss_scale == "dynamic" else float(loss_scale), dtype=acc_dtype, device=dev))
        self.scratch = torch.zeros(2, dtype=acc_dtype, device=dev)
        self.adam_state = torch.zeros(3, dtype=acc_dtype, device=dev)  # step, 1-b1^t, sqrt(1-b2^t) (device side)
        # the learning rate the step uses is lr0 * lr_scale[0], read on the device: set_lr() also reaches a captured graph
        self.lr0 = float(lr)
        self.lr_scale = torch.ones(1, dtype=acc_dtype, device=dev)
        if self.sync_bn and ops.option("p2p_syncbn") and self.acc_dtype == torch.float32 and dev.type == "cuda":
            ok = 1
            try:
                ex = ops.P2PExchange(self.group, dev)
            except Exception as e:  # noqa: BLE001 - no VMM / fabric support on this box, old torch, ...
                import sys
                print(f"gdl_b200: NVLink peer exchange unavailable ({type(e).__name__}: {e}); SyncBN statistics go through NCCL",
                      file=sys.stderr)
                ok, ex = 0, None
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 1:
                self.bn_exchange, self.bn_exchange_kind = ex, "nvlink_p2p"
        self.last_engine: Engine | None = None
        # CUDA graph of the whole step (normalise .. Adam): removes the ~10 us/launch host cost of the
        # ~800-2000 launches of a step.  Captured lazily on the first step() with a given input shape.
        self.cuda_graph = cuda_graph
        self.repack_in_place = True  # False: drop the packed-weight cache after every step (round-1 behaviour)
        self._graph = None
        self._static = None
        self._static_aug = None
        self.launches_per_step = 0

    @torch.no_grad()
    def forward_backward(self, image_u8: torch.Tensor, target: torch.Tensor,
                         aug_params: torch.Tensor | None = None) -> torch.Tensor:
        """image_u8: (N,H,W,C) uint8 on the device ((N,C,H,W) with input_chw); target: (N,H,W) int64/uint8. Returns the loss (device scalar)
        with d(loss)/d(params) left in the flat gradient buffer.  aug_params: optional int32 (N,6) device table of
        gdl_b200.augment.BatchAugmenter.sample(): the augmentation is then applied by the normalisation pass itself."""
        model = self.model
        for work in self._pending:  # bucket all-reduces of a backward whose optimizer_step was never called
            work.wait()
        self._pending = []
        self._next_bucket = 0  # every route: optimizer_step reduces the buckets the backward has not started itself
        self.gflat.zero_()
        eng = Engine(model.compute_dtype, training=True, wcache=model._wcache, grad_dst=self.grad_dst,
                     sync_bn_group=self.group if self.sync_bn else None, acc_dtype=self.acc_dtype,
                     sync_bn_exchange=self.bn_exchange)
        chw = self.input_chw
        c = image_u8.shape[1 if chw else 3]
        if aug_params is not None:
            x, target = ops.augment_normalize(image_u8, chw, target, aug_params, model.compute_dtype,
                                              (c + 7) // 8 * 8, self.mean, self.std, self.image_max)
        else:
            x = ops.normalize_to_nhwc(image_u8, chw, model.compute_dtype, (c + 7) // 8 * 8, self.mean, self.std,
                                      self.image_max)
        if hasattr(model, "fused_train"):
            # models with several logit maps / a frozen front half (DOFA + UperNet) own the whole step
            loss = (model.fused_train(eng, x, c, target, self.loss) if self.loss_scale is None
                    else model.fused_train(eng, x, c, target, self.loss, grad_scale=self.loss_scale))
            self.last_engine = eng
            return loss
        logits = (model.run(eng, Act(x, needs_grad=False), c) if getattr(model, "needs_bands", False)
                  else model.run(eng, Act(x, needs_grad=False)))
        coeff, _ = ops.seg_loss_fwd(logits, target, self.loss)
        n, h, w, k = logits.shape
        if hasattr(model, "backward"):
            # models that own their backward (SegFormer: the logits pass through a bilinear x4 first)
            d = torch.empty_like(logits)
            ops.seg_loss_bwd(logits, target, self.loss, coeff, self.loss_scale, d)
            model.backward(eng, d)
        else:
            # UNet++: the loss kernel writes the 16-bit, 16-channel-padded operand of the head's dgrad/wgrad
            d16 = torch.zeros((n, h, w, (k + 15) // 16 * 16), dtype=model.compute_dtype, device=logits.device)
            ops.seg_loss_bwd(logits, target, self.loss, coeff, self.loss_scale, d16)
            eng.head_backward(d16)
            if self.overlap_allreduce:
                eng.on_progress = self._reduce_finished_buckets
            eng.backward()
        self.last_engine = eng
        return coeff[0]

    def _reduce_finished_buckets(self, eng: Engine) -> None:
        """after a backward closure: start the all-reduce of every bucket (in backward order) whose parameters all have their
        gradient — the kernels that wrote it are already enqueued on this stream, NCCL's stream waits for them"""
        while self._next_bucket < len(self._buckets):
            a, b, ids = self._buckets[self._next_bucket]
            if not ids <= eng.param_grads.keys():
                return
            self._pending.append(dist.all_reduce(self.gflat[a:b], group=self.group, async_op=True))
            self._next_bucket += 1

    @torch.no_grad()
    def optimizer_step(self) -> None:
        if self.world > 1:
            if self.overlap_allreduce and self._buckets:
                for work in self._pending:
                    work.wait()  # the current stream waits for NCCL's
                for a, b, _ in self._buckets[self._next_bucket:]:  # whatever the backward did not finish early
                    dist.all_reduce(self.gflat[a:b], group=self.group)
                self._pending, self._next_bucket = [], len(self._buckets)
            else:
                dist.all_reduce(self.gflat, group=self.group)
            scale_by = 1.0 / self.world
        else:
            scale_by = None
        if self.loss_scale is not None:
            # unscale first: the clip threshold and Adam see true gradients
            self.gflat.mul_((1.0 / self.loss_scale).to(self.gflat.dtype))
            if self.loss_scale_mode == "dynamic":
                if not bool(torch.isfinite(self.gflat).all()):  # host decision, as GradScaler.step()
                    self.loss_scale.mul_(0.5)
                    self._good_steps = 0
                    self.skipped_steps += 1
                    return
                self._good_steps += 1
                if self._good_steps >= self.growth_interval:
                    self.loss_scale.mul_(2.0)
                    self._good_steps = 0
        gs = None
        if self.clip is not None or scale_by is not None:
            gs = self.scratch[1:2]
            if self.clip is not None:
                if scale_by is not None:
                    self.gflat.mul_(scale_by)
                ops.grad_clip_coef(self.gflat, self.clip, self.scratch[0:1], gs)
            else:
                gs.fill_(scale_by)
        self.step_count += 1
        ops.adam_step_dev(self.flat, self.gflat, self.m, self.v, self.lr0, self.betas[0], self.betas[1], self.eps,
                          self.weight_decay, self.adam_state, gs, self.lr_scale)
        # parameters changed behind torch's version counters: refresh the cached 16-bit operands in place (one launch)
        if not self.repack_in_place or not refresh_packed_weights(self.model._wcache):
            self.model._wcache.clear()

    def close(self) -> None:
        """Drop what lives in the process group (the NVLink exchange buffers, pending collectives).  Call it — or delete the
        trainer — BEFORE torch.distributed.destroy_process_group(): symmetric memory torn down after the group is gone blocks."""
        for work in self._pending:
            work.wait()
        self._pending = []
        self.bn_exchange = None
        self._graph = None

    def set_lr(self, lr: float) -> None:
        """Learning rate of the following steps (a scheduler's hook: ReduceLROnPlateau / OneCycleLR in the reference's
        YAMLs).  Stored on the device, so eager steps and replays of an already captured graph both use it."""
        self.lr = float(lr)
        self.lr_scale.fill_(self.lr / self.lr0 if self.lr0 != 0.0 else 0.0)

    def _eager_step(self, image_u8: torch.Tensor, target: torch.Tensor,
                    aug_params: torch.Tensor | None = None) -> torch.Tensor:
        loss = self.forward_backward(image_u8, target, aug_params)
        self.optimizer_step()
        return loss

    @torch.no_grad()
    def step(self, image_u8: torch.Tensor, target: torch.Tensor, aug_params: torch.Tensor | None = None) -> torch.Tensor:
        """One training step.  With cuda_graph=True the first call runs eagerly (lazy CUDA init, function
        attributes), the second captures, later calls copy the inputs into the static buffers and replay.
        aug_params (optional, int32 (N,6) on the device) turns the normalisation pass into augment + normalise; a
        trainer that captured its graph with / without it re-captures when that changes."""
        if not self.cuda_graph:
            n0 = ops.launch_count()
            loss = self._eager_step(image_u8, target, aug_params)
            self.launches_per_step = ops.launch_count() - n0
            return loss
        key = (tuple(image_u8.shape), image_u8.dtype, tuple(target.shape), target.dtype, aug_params is not None)
        if self._static is None or self._static[0] != key:
            self._static = (key, torch.empty_like(image_u8), torch.empty_like(target), 0)
            self._static_aug = torch.empty_like(aug_params) if aug_params is not None else None
            self._graph = None
        _, s_img, s_tgt, seen = self._static
        s_img.copy_(image_u8, non_blocking=True)
        s_tgt.copy_(target, non_blocking=True)
        s_aug = self._static_aug
        if s_aug is not None:
            s_aug.copy_(aug_params, non_blocking=True)
        if seen == 0:
            self._static = (key, s_img, s_tgt, 1)
            return self._eager_step(s_img, s_tgt, s_aug).clone()
        if self._graph is None:
            self.last_engine = None
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            if self.world > 1:
                # the NCCL collectives (flat-gradient all-reduce, SyncBN sums) are captured with the kernels; every rank
                # captures the same sequence.  Ranks must agree on capture vs eager, so a failure here is fatal.
                dist.barrier(group=self.group)
            with torch.cuda.graph(graph):
                self._graph_loss = self._eager_step(s_img, s_tgt, s_aug)
            self.launches_per_step = ops.launch_count() - n0
            self.last_engine = None  # activations live in the graph's private pool
            self._graph = graph
        self._graph.replay()
        return self._graph_loss
