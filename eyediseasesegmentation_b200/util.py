"""Host helpers on the inference path -- mirror of ``src/main/util/base_utils.py`` (the
names ``tta.py`` imports at line 34: ``lesion_dict, get_datapath, make_grid, multigen,
save_output``).  Pure host logic; tile order and coordinates are bit-compatible with the
reference because they decide which tile owns a pixel under overwrite-paste (tta.py:213).
"""
from __future__ import annotations

import collections
import os
import re
from pathlib import Path

import numpy as np

Lesion = collections.namedtuple("Lesion", ["dir_name", "project_name"])

# base_utils.py:22-41
lesion_dict = {
    "MA": Lesion("1. Microaneurysms", "MicroaneurysmsSegmentation"),
    "EX": Lesion("3. Hard Exudates", "HardExudatesSegmentation"),
    "HE": Lesion("2. Haemorrhages", "HaemorrhageSegmentation"),
    "SE": Lesion("4. Soft Exudates", "SoftExudatesSegmentation"),
    "MA_DDR": Lesion("MA", "DDRMicroaneurysmsSegmentation"),
    "EX_DDR": Lesion("EX", "DDRHardExudatesSegmentation"),
    "HE_DDR": Lesion("HE", "DDRHaemorrhageSegmentation"),
    "SE_DDR": Lesion("SE", "DDRSoftExudatesSegmentation"),
    "OD": Lesion("5. Optic Disc", "OpticDiscSegmentation"),
    "EX_FGADR": Lesion("HardExudate_Masks", "FGADRHardExudatesSegmentation"),
    "HE_FGADR": Lesion("Hemohedge_Masks", "FGADRHaemorrhageSegmentation"),
    "SE_FGADR": Lesion("SoftExudate_Masks", "FGADRSoftExudatesSegmentation"),
    "MA_FGADR": Lesion("Microaneurysms_Masks", "FGADRMicroaneurysmsSegmentation"),
    "Vessel_DRIVE": Lesion("", "DRIVE_VesselSegmentation"),
    "Vessel_HRF": Lesion("", "HRF_VesselSegmentation"),
    "Vessel_CHASEDB1": Lesion("", "CHASEDB1_VesselSegmentation"),
}


def multigen(gen_func):
    """Decorator that makes a generator function re-iterable (base_utils.py:43-50)."""

    class _Reiterable:
        def __init__(self, *args, **kwargs):
            self._args, self._kwargs = args, kwargs

        def __iter__(self):
            return gen_func(*self._args, **self._kwargs)

    return _Reiterable


def _starts(size: int, window: int, min_overlap: int) -> np.ndarray:
    count = size // (window - min_overlap) + 1
    starts = np.linspace(0, size, num=count, endpoint=False, dtype=np.int64)
    starts[-1] = size - window
    return starts


def make_grid(shape, window=256, min_overlap=32):
    """(N, 4) int64 array of tile slices ``x1, x2, y1, y2`` (rows first), row-major over the
    tile grid -- same values as base_utils.py:52-71, including its degenerate cases
    (negative starts when window > image, repeated tiles when window == image)."""
    rows, cols = shape
    r1 = _starts(rows, window, min_overlap)
    c1 = _starts(cols, window, min_overlap)
    r2 = (r1 + window).clip(0, rows)
    c2 = (c1 + window).clip(0, cols)
    out = np.zeros((len(r1) * len(c1), 4), dtype=np.int64)
    k = 0
    for i in range(len(r1)):
        for j in range(len(c1)):
            out[k] = (r1[i], r2[i], c1[j], c2[j])
            k += 1
    return out


def get_datapath(img_path, mask_path, lesion_type: str = "EX"):
    """Sorted (image paths, mask paths) for a dataset layout (base_utils.py:82-122)."""
    parts = lesion_type.split("_")
    if parts[0] == "Vessel":
        return sorted(img_path.glob("*.jpg")), sorted(mask_path.glob("*.jpg"))
    if len(parts) == 1:
        sub = lesion_dict[lesion_type].dir_name
        suffix = "_" + lesion_type + ".tif"
        mask_names = os.listdir(os.path.join(mask_path, sub))
        images = [Path(os.path.join(img_path, re.sub(suffix, "", m) + ".jpg")) for m in mask_names]
        masks = [Path(os.path.join(mask_path, sub, m)) for m in mask_names]
        return sorted(images), sorted(masks)
    sub = lesion_dict[lesion_type].dir_name
    if parts[1] == "FGADR":
        return sorted(img_path.glob("*.png")), sorted((mask_path / sub).glob("*.png"))
    if parts[1] == "DDR":
        if isinstance(img_path, tuple):
            imgs = tuple(sorted(p.glob("*.jpg")) for p in img_path[:2])
            masks = tuple(sorted((p / sub).glob("*.tif")) for p in mask_path[:2])
            return imgs, masks
        return sorted(img_path.glob("*.jpg")), sorted((mask_path / sub).glob("*.tif"))
    raise KeyError(lesion_type)


def save_output(pred_masks: np.ndarray, out_path: Path):
    """Min-max rescale to 0..255 uint8 and save with PIL (base_utils.py:124-131)."""
    from PIL import Image
    pred_masks = np.asarray(pred_masks)
    scaled = (255.0 / (pred_masks.max() + np.finfo(float).eps) * (pred_masks - pred_masks.min())).astype(np.uint8)
    Image.fromarray(scaled).save(out_path)
    print(f"[INFO] saved {Path(out_path).name} to disk")
