"""Mirror of ``src/data/augment_vessel/pad_img.py`` (the offline step that produces the square, /32
inputs ``tta_vessel.test_tta`` expects: DRIVE 584x565 -> 608^2, CHASEDB1 960x999 -> 1024^2).

``pad(input_folder, output_folder, desired_size=608, already_padded=False, is_mask=False)`` keeps the
reference's signature and file behaviour (pad_img.py:8-38): every file of ``input_folder`` (sorted) is
centre-padded with zeros to ``desired_size`` -- ``top = dh // 2``, ``bottom = dh - dh // 2``, same for
left / right -- and written under the same name; masks are read as single-channel grey and binarised
with ``> 127 -> 255`` after padding.  ``pad_array`` is the same arithmetic on an array, used by the
drivers when they are handed unpadded vessel images (``vessel_pad_geometry``).

Host-side byte shuffling (no arithmetic on the GPU path); cv2 does the file I/O as in the reference.
The reference reads masks with ``skimage.io.imread(as_gray=True)``, which returns a 2-D uint8 array
unchanged for single-channel files (the DRIVE / CHASEDB1 label format); PIL's ``L`` mode read is the
same bytes.
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["pad", "pad_array", "pad_geometry"]


def pad_geometry(shape_hw, desired_size: int):
    """(top, bottom, left, right) of pad_img.py:21-25."""
    h, w = int(shape_hw[0]), int(shape_hw[1])
    delta_w = desired_size - w
    delta_h = desired_size - h
    if delta_w < 0 or delta_h < 0:
        # cv2.copyMakeBorder raises on negative borders
        raise ValueError(f"image {h}x{w} is larger than desired_size {desired_size}")
    top, bottom = delta_h // 2, delta_h - (delta_h // 2)
    left, right = delta_w // 2, delta_w - (delta_w // 2)
    return top, bottom, left, right


def pad_array(img: np.ndarray, desired_size: int = 608, is_mask: bool = False) -> np.ndarray:
    """Centre zero pad to ``desired_size`` x ``desired_size``; masks are then thresholded ``> 127 -> 255``."""
    top, bottom, left, right = pad_geometry(img.shape[:2], desired_size)
    out = np.zeros((desired_size, desired_size) + img.shape[2:], dtype=img.dtype)
    out[top:top + img.shape[0], left:left + img.shape[1]] = img
    if is_mask:
        out = np.where(out > 127, 255, 0).astype(img.dtype)
    return out


def pad(input_folder, output_folder, desired_size=608, already_padded=False, is_mask=False):
    import cv2
    from PIL import Image
    if not os.path.exists(output_folder):
        os.makedirs(output_folder)
    for file in sorted(os.listdir(input_folder)):
        print(file)
        path = str(input_folder) + "/" + file
        if is_mask:
            tmp = np.asarray(Image.open(path).convert("L"))
        else:
            tmp = cv2.imread(path)              # BGR in, BGR out: the file keeps its colours
        if not already_padded:
            tmp = pad_array(tmp, desired_size, is_mask)
        cv2.imwrite(str(output_folder) + "/" + file, tmp)
    if not already_padded:
        print("Padding is done.")
