"""TEST INFRASTRUCTURE: restatement of the reference's vessel padding
(src/data/augment_vessel/pad_img.py:8-38) with the same OpenCV calls the reference makes.

  pad_array   :19-33  copyMakeBorder(top = dh // 2, bottom = dh - dh // 2, left = dw // 2, right = dw - dw // 2,
                      BORDER_CONSTANT 0) and, for masks, threshold(127, 255, THRESH_BINARY)

Pinned to the reference's own ``pad`` (run file-to-file through oracle/ref_loader.py) by
tests/golden/pad_img.npz (tests/golden/make_golden.py) and, in the build container, directly.
"""
from __future__ import annotations

import cv2
import numpy as np


def pad_array(img: np.ndarray, desired_size: int, is_mask: bool) -> np.ndarray:
    old_size = img.shape[:2]
    delta_w = desired_size - old_size[1]
    delta_h = desired_size - old_size[0]
    top, bottom = delta_h // 2, delta_h - (delta_h // 2)
    left, right = delta_w // 2, delta_w - (delta_w // 2)
    color = [0] if is_mask else [0, 0, 0]
    out = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    if is_mask:
        _, out = cv2.threshold(out, 127, 255, cv2.THRESH_BINARY)
    return out
