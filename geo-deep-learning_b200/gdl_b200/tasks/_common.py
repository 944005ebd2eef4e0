"""What the three task mirrors share with the reference's tasks outside the hot path: checkpoint loading
(`utils/models.py:12-66` load_weights_from_checkpoint: optional `load_parts` filter + strict=False), the optimizer /
scheduler wiring of `configure_optimizers` (tasks_with_models/segmentation_segformer.py:150-199 and its UNet++ / DOFA
twins: OneCycleLR gets its horizon from the trainer), and the handling of the `weights` hyper-parameter."""
from __future__ import annotations

import logging
import math
from typing import Any

import torch
from torch import Tensor

logger = logging.getLogger(__name__)


def strip_model_prefix(sd: dict[str, Tensor]) -> dict[str, Tensor]:
    return {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}


def load_weights_from_checkpoint(model: torch.nn.Module, path: str, load_parts: str | list[str] | None = None,
                                 map_location: Any = "cpu", trust_pickle: bool = False):
    """utils/models.py:12-66.  torch.load runs with its default weights_only=True (as on the reference's pinned torch);
    `trust_pickle=True` opts out for checkpoints that hold arbitrary Python objects — only for files you trust."""
    logger.info("Loading weights from checkpoint: %s", path)
    ckpt = torch.load(path, map_location=map_location, weights_only=not trust_pickle)
    sd = strip_model_prefix(ckpt.get("state_dict", ckpt))
    if load_parts is None:
        model.load_state_dict(sd)
        return None
    parts = [load_parts] if isinstance(load_parts, str) else list(load_parts)
    sd = {k: v for k, v in sd.items() if any(k.startswith(f"{p}.") for p in parts)}
    result = model.load_state_dict(sd, strict=False)
    for p in parts:
        n = sum(k.startswith(f"{p}.") for k in sd)
        logger.info("  - %s: %s", p, f"{n} parameters loaded" if n else "NO PARAMETERS FOUND - check if this part exists")
    logger.info("Missing keys: %s, unexpected keys: %s", len(result.missing_keys), len(result.unexpected_keys))
    return result


def hparam(task, name: str, default=None):
    hp = task.hparams
    try:
        return hp.get(name, default)
    except AttributeError:
        return getattr(hp, name, default)


def check_pretrained_request(weights: Any, what: str) -> None:
    """The reference downloads ImageNet / Hugging Face weights when `weights` (or `pretrained`) is set (its YAMLs ship
    `weights: imagenet`).  There is no network here and the models refuse the request — so must the task, loudly, instead
    of silently training from random initialisation."""
    if weights is None or weights is False:
        return
    raise ValueError(
        f"{what}: weights={weights!r} asks for a pretrained download, which this build cannot do (no network). Either set "
        "weights: null and pass the pretrained checkpoint through `weights_from_checkpoint_path` (state_dict keys are the "
        "reference's), or load the tensors yourself with model.load_state_dict().")


def configure_optimizers(task):
    """segmentation_segformer.py:150-199 (identical in the UNet++ and DOFA tasks).  With a LightningCLI scheduler
    dictionary in the hyper-parameters, OneCycleLR gets its horizon from the trainer (estimated stepping batches, else
    the datamodule's epoch_size / batch_size, else the configured total_steps) and any other class is built by the
    `scheduler` callable; without one (direct construction) the callable is applied as is, and no scheduler is returned
    when there is none."""
    opt = task.optimizer(task.parameters())
    cfg = hparam(task, "scheduler")
    if not isinstance(cfg, dict):
        sched = task.scheduler(opt) if callable(task.scheduler) else None
        return ([opt], [{"scheduler": sched, **task.scheduler_config}]) if sched is not None else [opt]
    if cfg.get("class_path", "") == "torch.optim.lr_scheduler.OneCycleLR":
        init = cfg.get("init_args", {}) or {}
        max_lr = init.get("max_lr")
        stepping = task.trainer.estimated_stepping_batches
        dm = getattr(task.trainer, "datamodule", None)
        if stepping > -1:
            sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, total_steps=stepping)
        elif getattr(dm, "epoch_size", None) is not None:
            accum = task.trainer.accumulate_grad_batches
            per_epoch = math.ceil(dm.epoch_size / (dm.batch_size * accum))
            sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, steps_per_epoch=per_epoch + int(per_epoch * accum),
                                                        epochs=task.trainer.max_epochs)
        else:
            sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, total_steps=init.get("total_steps"))
    else:
        sched = task.scheduler(opt)
    return [opt], [{"scheduler": sched, **task.scheduler_config}]
