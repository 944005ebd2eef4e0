"""ORACLE (test infrastructure). Restatement of the per-sample path of the reference's WebDataset pipeline,
geo_deep_learning/datasets/wds_dataset.py:
  * `create_shard_split_paths` (:52-80), `_load_normalization_stats` (:198-215: mean / 255, std / 255, viewed (C,1,1));
  * `_process_sample` (:217-243): image = from_numpy(image_patch).float(); label = from_numpy(label_patch).long();
    `normalization` then `standardization` (utils/tensors.py), then the clay / dofa / unified dictionaries (:245-303);
  * `_encode_temporal` / `_encode_spatial` / `_extract_wavelengths` (:305-390).
Decoding of the shards themselves is webdataset's (tar members grouped by key, `.npy` -> numpy.load, `.json` ->
json.loads); the tests write their shards with numpy.save / json.dumps, so numpy itself is the decoder oracle.

PINNED: tests/golden/wds_golden.pt holds outputs of the REFERENCE's own `ShardedDataset._process_sample` (imported in the
build container by oracle/make_golden.py with `webdataset` / `pytorch_lightning` satisfied by empty stand-ins — neither is
touched by `_process_sample`) for the dofa, clay and unified formats.
"""
from __future__ import annotations

import json
import math
from datetime import datetime
from pathlib import Path
from typing import Any

import numpy as np
import torch

from . import tensors as ot


def create_shard_split_paths(manifest_path: str, split: str, parent_dir: str | None = None):
    parent = Path(manifest_path).parent / split if parent_dir is None else Path(parent_dir) / split
    data = json.loads(Path(manifest_path).read_text())
    return [(parent / it["path"]).as_posix() for it in data["shards"][split]], data["statistics"]["patch_counts"][split]


def load_normalization_stats(stats_path: str, sensor_name: str) -> dict[str, Any]:
    st = json.loads(Path(stats_path).read_text())["statistics"][sensor_name]
    return {"mean": torch.tensor(st["mean"], dtype=torch.float32).div(255.0).view(-1, 1, 1),
            "std": torch.tensor(st["std"], dtype=torch.float32).div(255.0).view(-1, 1, 1),
            "band_count": st["band_count"], "patch_count": st["patch_count"], "dtype": st["dtype"]}


def encode_temporal(s: str) -> torch.Tensor:
    try:
        if s.endswith("Z"):
            s = s[:-1] + "+00:00"
        dt = datetime.fromisoformat(s)
        wr, hr = (dt.isocalendar().week / 52.0) * 2 * math.pi, (dt.hour / 24.0) * 2 * math.pi
        return torch.tensor([math.sin(wr), math.cos(wr), math.sin(hr), math.cos(hr)], dtype=torch.float32)
    except Exception:  # noqa: BLE001
        return torch.zeros(4, dtype=torch.float32)


def encode_spatial(lat: float, lon: float) -> torch.Tensor:
    try:
        a, o = math.radians(lat), math.radians(lon)
        return torch.tensor([math.sin(a), math.cos(a), math.sin(o), math.cos(o)], dtype=torch.float32)
    except Exception:  # noqa: BLE001
        return torch.zeros(4, dtype=torch.float32)


def extract_wavelengths(metadata: dict[str, Any], keys: list[str] | None, sensor_name: str = "",
                        cache: dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """:357-390.  The reference keeps a per-dataset cache keyed by sensor + key names and returns the FIRST sample's
    wavelengths for every later sample of that sensor (:378-386); pass the dataset's `cache` dict to follow that."""
    keys = keys or ["red_wavelength", "green_wavelength", "blue_wavelength", "nir_wavelength"]
    try:
        meta = metadata["metadata"]
        w = [float(meta[b]) for b in keys if b in meta]
        if cache is None:
            return torch.tensor(w, dtype=torch.float32)
        ck = f"{sensor_name}_{'_'.join(keys)}"
        if ck not in cache:
            cache[ck] = torch.tensor(w, dtype=torch.float32)
        return cache[ck]
    except Exception:  # noqa: BLE001
        return torch.tensor([0.0] * len(keys), dtype=torch.float32)


def process_sample(sample: dict[str, Any], stats: dict[str, Any], sensor_name: str, model_type: str,
                   wavelength_keys: list[str] | None = None, wl_cache: dict[str, torch.Tensor] | None = None) -> dict[str, Any]:
    """sample: {"__key__", "image_patch.npy": ndarray, "label_patch.npy": ndarray, "metadata.json": dict}"""
    image = torch.from_numpy(np.ascontiguousarray(sample["image_patch.npy"])).float()
    label = torch.from_numpy(np.ascontiguousarray(sample["label_patch.npy"])).long()
    metadata = sample["metadata.json"]
    image = ot.normalization(image)
    image = ot.standardization(image, stats["mean"], stats["std"])
    out = {"image": image, "mask": label, "platform": sensor_name, "image_name": sample["__key__"],
           "mean": stats["mean"], "std": stats["std"]}
    if model_type == "clay":
        m = metadata["metadata"]
        out["time"] = encode_temporal(m.get("datetime", "0.0"))
        out["latlon"] = encode_spatial(m.get("coordinates_lat", 0.0), m.get("coordinates_lon", 0.0))
    elif model_type == "dofa":
        out["wavelengths"] = extract_wavelengths(metadata, wavelength_keys, sensor_name, wl_cache)
    else:
        out["metadata"] = metadata
    return out
