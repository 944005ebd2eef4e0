"""WebDataset shard -> GPU feeder: the data format in front of the patch normalisation.

Reference: `geo_deep_learning/datasets/wds_dataset.py` — tar shards of samples
`<key>.image_patch.npy` (C,H,W), `<key>.label_patch.npy` (1,H,W) or (H,W), `<key>.metadata.json`; a manifest JSON
(`shards[split][].path`, `statistics.patch_counts[split]`, :52-80); a statistics JSON
(`statistics[sensor].{mean,std,band_count,patch_count,dtype}`, mean/std in 0..255 units, :198-215).  There every sample
is decoded to float32 on a CPU worker, divided by 255, standardised (:230-236), batched and only then copied to the GPU
(4 bytes per value over PCIe).

Here the payload bytes of the `.npy` members are copied as they are into a pinned batch buffer (no per-sample tensors, no
float conversion on the host), copied as uint8 (1 byte per value) on a side stream while the previous batch trains, and normalised by the kernel
that feeds the stem (`gdl_normalize_to_nhwc`, in_kind = uint8 NCHW) — or by `gdl_augment_normalize` when the batch is also
augmented.  What is mirrored from the reference: shard list per split / rank (`shard_list[rank::world_size]` for "trn",
:398-401), per-sensor statistics divided by 255, the dofa / clay / unified sample dictionaries (:238-303), wavelength
extraction (:357-390), temporal / spatial encodings (:305-355), `partial` last batch for every split but "trn" (:418-422).
webdataset's own shuffling RNG cannot be reproduced (the package is not vendored): the shuffle buffer here is a seeded
restatement of the same algorithm (fill a buffer, emit a random element, refill).

No kernel is launched here apart from the optional normalisation; everything else is host IO.  The webdataset layout
needs no third-party package: tar members are walked with the standard library, `.npy` headers are parsed here.
"""
from __future__ import annotations

import ast
import json
import math
import random
import struct
import tarfile
import threading
import queue
from datetime import datetime
from pathlib import Path
from typing import Any, Iterator

import numpy as np
import torch

from . import ops

IMAGE_EXT, LABEL_EXT, META_EXT = "image_patch.npy", "label_patch.npy", "metadata.json"


# ---------------------------------------------------------------------------------------------
# manifest / statistics (wds_dataset.py:52-80, 198-215)
# ---------------------------------------------------------------------------------------------
def create_shard_split_paths(manifest_path: str, split: str, parent_dir: str | None = None) -> tuple[list[str], int]:
    base = (Path(manifest_path).parent if parent_dir is None else Path(parent_dir)) / split
    with Path(manifest_path).open() as f:
        data = json.load(f)
    return ([(base / item["path"]).as_posix() for item in data["shards"][split]],
            data["statistics"]["patch_counts"][split])


def load_normalization_stats(stats_path: str, sensor_name: str) -> dict[str, Any]:
    with Path(stats_path).open() as f:
        stats = json.load(f)["statistics"][sensor_name]
    return {"mean": torch.tensor(stats["mean"], dtype=torch.float32).div(255.0),
            "std": torch.tensor(stats["std"], dtype=torch.float32).div(255.0),
            "band_count": stats["band_count"], "patch_count": stats["patch_count"], "dtype": stats["dtype"]}


# ---------------------------------------------------------------------------------------------
# .npy header (numpy format 1.0 / 2.0 / 3.0) — enough to locate the raw payload inside the tar member
# ---------------------------------------------------------------------------------------------
def parse_npy_header(head: bytes) -> tuple[np.dtype, tuple[int, ...], bool, int]:
    """-> (dtype, shape, fortran_order, payload offset).  `head` = the first bytes of the member (>= header length)."""
    if len(head) < 10 or head[:6] != b"\x93NUMPY":
        raise ValueError("not a .npy member (bad magic)")
    major = head[6]
    if major == 1:
        (hlen,) = struct.unpack("<H", head[8:10])
        start = 10
    elif major in (2, 3):
        if len(head) < 12:
            raise ValueError("truncated .npy header")
        (hlen,) = struct.unpack("<I", head[8:12])
        start = 12
    else:
        raise ValueError(f"unsupported .npy format version {major}")
    if len(head) < start + hlen:
        raise ValueError("truncated .npy header")
    d = ast.literal_eval(head[start:start + hlen].decode("latin1" if major < 3 else "utf8"))
    dtype = np.dtype(d["descr"])
    if dtype.hasobject:
        raise ValueError("object arrays are not supported")
    return dtype, tuple(int(v) for v in d["shape"]), bool(d["fortran_order"]), start + hlen


def _split_key(name: str) -> tuple[str, str]:
    """webdataset's base_plus_ext: the key ends at the first dot of the file name; the rest is the extension."""
    slash = name.rfind("/")
    dot = name.find(".", slash + 1)
    return (name, "") if dot < 0 else (name[:dot], name[dot + 1:])


def iter_tar_samples(path: str) -> Iterator[dict[str, Any]]:
    """Yields {"__key__": key, ext: bytes, ...} for each group of consecutive members sharing a key."""
    with tarfile.open(path, "r:*") as tf:
        cur: dict[str, Any] | None = None
        for m in tf:
            if not m.isfile():
                continue
            key, ext = _split_key(m.name)
            if not ext:
                continue
            if cur is None or cur["__key__"] != key:
                if cur is not None:
                    yield cur
                cur = {"__key__": key}
            cur[ext] = tf.extractfile(m).read()
        if cur is not None:
            yield cur


def decode_npy(buf: bytes) -> np.ndarray:
    dtype, shape, fortran, off = parse_npy_header(buf[:4096] if len(buf) > 4096 else buf)
    n = int(np.prod(shape)) if shape else 1
    a = np.frombuffer(buf, dtype=dtype, count=n, offset=off)
    return a.reshape(shape, order="F" if fortran else "C")


# ---------------------------------------------------------------------------------------------
# metadata encodings (wds_dataset.py:305-390)
# ---------------------------------------------------------------------------------------------
def encode_temporal(datetime_str: str) -> torch.Tensor:
    try:
        if datetime_str.endswith("Z"):
            datetime_str = datetime_str[:-1] + "+00:00"
        dt = datetime.fromisoformat(datetime_str)
        week_rad = (dt.isocalendar().week / 52.0) * 2 * math.pi
        hour_rad = (dt.hour / 24.0) * 2 * math.pi
        return torch.tensor([math.sin(week_rad), math.cos(week_rad), math.sin(hour_rad), math.cos(hour_rad)],
                            dtype=torch.float32)
    except Exception:  # noqa: BLE001 - the reference logs and returns zeros
        return torch.zeros(4, dtype=torch.float32)


def encode_spatial(lat: float, lon: float) -> torch.Tensor:
    try:
        la, lo = math.radians(lat), math.radians(lon)
        return torch.tensor([math.sin(la), math.cos(la), math.sin(lo), math.cos(lo)], dtype=torch.float32)
    except Exception:  # noqa: BLE001
        return torch.zeros(4, dtype=torch.float32)


DEFAULT_WAVELENGTH_KEYS = ["red_wavelength", "green_wavelength", "blue_wavelength", "nir_wavelength"]


def extract_wavelengths(metadata: dict[str, Any], wavelength_keys: list[str] | None) -> torch.Tensor:
    keys = wavelength_keys or DEFAULT_WAVELENGTH_KEYS
    try:
        meta = metadata["metadata"]
        return torch.tensor([float(meta[k]) for k in keys if k in meta], dtype=torch.float32)
    except Exception:  # noqa: BLE001
        return torch.tensor([0.0] * len(keys), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------
# the feeder
# ---------------------------------------------------------------------------------------------
class ShardFeeder:
    """Iterates batches of one sensor / split.  Each batch is a dict:

        image_u8    (B,C,H,W) uint8 on `device` (raw payload; f32 when the shard does not hold uint8)
        mask        (B, *label_patch.shape) int64 on `device` — (B,1,H,W) for the reference's (1,H,W) labels
                    (uint8 with mask_dtype=torch.uint8: 8x fewer bytes over PCIe)
        mean, std   (C,) float32 on `device` (statistics / 255)
        platform, image_name (list of keys), and per model_type: wavelengths | time, latlon | metadata
        image       only with normalize=True: f32 (B,C,H,W), ((x / 255) - mean) / std — the reference's batch["image"]

    `FusedTrainer(..., input_chw=True).step(batch["image_u8"], batch["mask"][:, 0].contiguous())` consumes the raw
    form directly."""

    def __init__(self, sensor_name: str, shard_paths: list[str], stats: dict[str, Any], *, model_type: str = "clay",
                 split: str = "trn", batch_size: int = 16, shuffle_buffer: int = 0, shardshuffle: int | None = None,
                 seed: int = 42, wavelength_keys: list[str] | None = None, device: torch.device | str = "cuda",
                 rank: int | None = None, world_size: int | None = None, depth: int = 2, normalize: bool = False,
                 mask_dtype: torch.dtype = torch.int64) -> None:
        if model_type not in ("clay", "dofa", "unified"):
            raise ValueError(f"unknown model_type {model_type!r}")
        if split not in ("trn", "val", "tst"):
            raise ValueError(f"unknown split {split!r}")
        if mask_dtype not in (torch.int64, torch.uint8):
            raise ValueError("mask_dtype must be int64 or uint8")
        self.sensor_name, self.model_type, self.split = sensor_name, model_type, split
        self.batch_size, self.shuffle_buffer, self.shardshuffle, self.seed = batch_size, shuffle_buffer, shardshuffle, seed
        self.wavelength_keys = wavelength_keys
        self.device = torch.device(device)
        self.depth, self.normalize, self.mask_dtype = max(1, depth), normalize, mask_dtype
        self.stats = stats
        self.epoch = 0
        if rank is None or world_size is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank, world_size = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        shards = sorted(shard_paths)
        if split == "trn" and world_size > 1:
            shards = shards[rank::world_size]  # wds_dataset.py:398-401
        self.shards = shards
        self._wl_cache: dict[str, torch.Tensor] = {}

    # -- host side: samples ----------------------------------------------------------------------
    def _samples(self) -> Iterator[dict[str, Any]]:
        shards = list(self.shards)
        rng = random.Random(self.seed + self.epoch)
        if self.split == "trn" and self.shardshuffle:
            rng.shuffle(shards)
        stream = (s for p in shards for s in iter_tar_samples(p))
        if self.split != "trn" or self.shuffle_buffer <= 1:
            yield from stream
            return
        buf: list[dict[str, Any]] = []
        for s in stream:  # webdataset's shuffle: keep `shuffle_buffer` samples, emit a random one per arrival
            if len(buf) < self.shuffle_buffer:
                buf.append(s)
                continue
            i = rng.randrange(len(buf))
            out, buf[i] = buf[i], s
            yield out
        rng.shuffle(buf)
        yield from buf

    def _extras(self, sample: dict[str, Any]) -> dict[str, Any]:
        metadata = json.loads(sample[META_EXT]) if META_EXT in sample else {"metadata": {}}
        if self.model_type == "clay":
            meta = metadata["metadata"]
            return {"time": encode_temporal(meta.get("datetime", "0.0")),
                    "latlon": encode_spatial(meta.get("coordinates_lat", 0.0), meta.get("coordinates_lon", 0.0))}
        if self.model_type == "dofa":
            keys = self.wavelength_keys or DEFAULT_WAVELENGTH_KEYS
            ck = f"{self.sensor_name}_{'_'.join(keys)}"
            try:  # the reference caches the first sample's wavelengths per sensor and returns them from then on (:378-386)
                w = [float(metadata["metadata"][k]) for k in keys if k in metadata["metadata"]]
                if ck not in self._wl_cache:
                    self._wl_cache[ck] = torch.tensor(w, dtype=torch.float32)
                return {"wavelengths": self._wl_cache[ck]}
            except Exception:  # noqa: BLE001 - logged and replaced by zeros in the reference (:388-390)
                return {"wavelengths": torch.tensor([0.0] * len(keys), dtype=torch.float32)}
        return {"metadata": metadata}

    def _host_batches(self) -> Iterator[dict[str, Any]]:
        """Batches staged in (pinned) host memory: the payload of every member is copied as raw bytes into the batch
        buffer (no per-sample tensors, no float conversion for uint8 shards)."""
        pin = self.device.type == "cuda"
        pending: list[dict[str, Any]] = []

        def flush(items: list[dict[str, Any]]) -> dict[str, Any]:
            imgs = [decode_npy(s[IMAGE_EXT]) for s in items]
            lbls = [decode_npy(s[LABEL_EXT]) for s in items]
            shape, dt = imgs[0].shape, imgs[0].dtype
            if any(a.shape != shape for a in imgs):
                raise ValueError("all patches of a batch must share one shape")
            raw = dt == np.uint8
            image = torch.empty((len(items), *shape), dtype=torch.uint8 if raw else torch.float32, pin_memory=pin)
            inp = image.numpy()
            for i, a in enumerate(imgs):
                inp[i] = a  # uint8: a byte copy; anything else: numpy's cast to float32 == torch's .float()
            mask = torch.empty((len(items), *lbls[0].shape), dtype=self.mask_dtype, pin_memory=pin)
            mnp = mask.numpy()
            for i, a in enumerate(lbls):
                mnp[i] = a  # numpy's integer cast == torch's .long()
            batch: dict[str, Any] = {"image_u8": image, "mask": mask, "platform": [self.sensor_name] * len(items),
                                     "image_name": [s["__key__"] for s in items]}
            extras = [self._extras(s) for s in items]
            for k in extras[0]:
                vals = [e[k] for e in extras]
                batch[k] = torch.stack(vals) if isinstance(vals[0], torch.Tensor) else vals
            return batch

        for s in self._samples():
            if IMAGE_EXT not in s or LABEL_EXT not in s:
                continue  # handler=wds.warn_and_continue
            pending.append(s)
            if len(pending) == self.batch_size:
                yield flush(pending)
                pending = []
        if pending and self.split != "trn":  # .batched(partial=self.split != "trn")
            yield flush(pending)

    # -- device side ---------------------------------------------------------------------------
    def _to_device(self, batch: dict[str, Any], stream) -> dict[str, Any]:
        dev = self.device
        out = dict(batch)
        if dev.type == "cuda":
            with torch.cuda.stream(stream):
                for k, v in batch.items():
                    if isinstance(v, torch.Tensor):
                        out[k] = v.to(dev, non_blocking=True)
                out["mean"] = self._mean_dev
                out["std"] = self._std_dev
                ev = torch.cuda.Event()
                ev.record(stream)
            out["_ready"] = ev
            out["_host"] = batch  # keeps the pinned buffers alive until the copies are done
        else:
            out["mean"], out["std"] = self._mean_dev, self._std_dev
        return out

    def __iter__(self) -> Iterator[dict[str, Any]]:
        dev = self.device
        self._mean_dev = self.stats["mean"].to(dev)
        self._std_dev = self.stats["std"].to(dev)
        stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        q: queue.Queue = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def producer() -> None:
            try:
                for hb in self._host_batches():
                    if stop.is_set():
                        return
                    q.put(self._to_device(hb, stream))
                q.put(None)
            except BaseException as e:  # noqa: BLE001 - re-raised in the consumer
                q.put(e)

        t = threading.Thread(target=producer, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                ev = item.pop("_ready", None)
                item.pop("_host", None)
                if ev is not None:
                    torch.cuda.current_stream(dev).wait_event(ev)
                    for v in item.values():
                        if isinstance(v, torch.Tensor) and v.is_cuda:
                            v.record_stream(torch.cuda.current_stream(dev))
                if self.normalize:
                    item["image"] = self.normalized_image(item)
                yield item
        finally:
            stop.set()
            while t.is_alive():  # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)
            self.epoch += 1

    def normalized_image(self, batch: dict[str, Any]) -> torch.Tensor:
        """f32 (B,C,H,W) = ((x / 255) - mean) / std on the device — the reference's batch["image"] (:230-236)."""
        x = batch["image_u8"]
        ident = torch.zeros((x.shape[0], 6), dtype=torch.int32, device=x.device)
        img, _ = ops.augment_normalize(x.contiguous(), True, None, ident, torch.float32, 0, batch["mean"], batch["std"], 255.0)
        return img
