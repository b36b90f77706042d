"""TEST INFRASTRUCTURE: CPU restatement of the reference's per-image inference loops.

  make_grid                  src/main/util/base_utils.py:52-71
  preprocessing              src/main/archs/__init__.py:88-97 (float64 numpy)
  tiled_probability_map      src/main/tta.py:196-213 (window -> cv2 resize -> normalise ->
                             TTA net -> sigmoid -> cv2 bilinear x2 -> overwrite paste)
  whole_image_probability    src/main/tta.py:108-121 (center crop + cv2 resize to full size)
  ensemble_probability       ensemble.py:86-100 (per-model sigmoid of the D4-mean logits, summed in
                             model order, divided by the model count).  Pinned by tests/golden/ensemble.npz,
                             written by the reference's own ensemble.py (ref_loader.load_ensemble runs it
                             with the absent third-party packages restated): tests/test_oracle.py

Pinned: tests/golden/tta_patches.npz and tta_whole.npz were written by the reference's own tta.py (run unmodified
through oracle/ref_loader.load_tta by tests/golden/make_golden.py); tests/test_oracle.py holds
tiled_probability_map and whole_image_probability to them (2e-6, exact in the build container).

``net`` is any callable ``[B,3,S,S] float32 tensor -> logits [B,1,S,S]`` (the oracle nets, or
the reference modules in the build container).  cv2 is the same library the reference calls.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch

from .nets import tta_mean_logits

DATASET_STATS = {
    "IDRiD": ([0.44976714, 0.2186806, 0.06459363], [0.33224553, 0.17116262, 0.086509705]),
    "DRIVE": ([0.49742976, 0.27066445, 0.16217253], [0.34794736, 0.18998094, 0.1084089]),
    "CHASEDB1": ([0.4527923, 0.16221291, 0.028265305], [0.36041078, 0.14167951, 0.036878455]),
}


def make_grid(shape, window=256, min_overlap=32):
    x, y = shape
    nx = x // (window - min_overlap) + 1
    x1 = np.linspace(0, x, num=nx, endpoint=False, dtype=np.int64)
    x1[-1] = x - window
    x2 = (x1 + window).clip(0, x)
    ny = y // (window - min_overlap) + 1
    y1 = np.linspace(0, y, num=ny, endpoint=False, dtype=np.int64)
    y1[-1] = y - window
    y2 = (y1 + window).clip(0, y)
    return np.array([[x1[i], x2[i], y1[j], y2[j]] for i in range(nx) for j in range(ny)], dtype=np.int64)


def preprocess(img_u8: np.ndarray, mean, std) -> np.ndarray:
    x = img_u8 / 255.0
    x = x - np.array(mean)
    x = x / np.array(std)
    return x


def gaussian_window(S2: int, sigma_scale: float = 0.25) -> np.ndarray:
    """Blend window of the product's OPT-IN Gaussian mode (no reference counterpart: the reference overwrites,
    tta.py:213): g[t] = exp(-(t - c)^2 / (2 sigma^2)), c = (S2 - 1) / 2, sigma = sigma_scale * S2."""
    t = np.arange(S2, dtype=np.float64)
    return np.exp(-((t - (S2 - 1) / 2.0) ** 2) / (2.0 * (sigma_scale * S2) ** 2)).astype(np.float32)


def tiled_probability_map(image_u8: np.ndarray, net, S: int, mean, std, tta_kind: str, blend: str = "overwrite") -> np.ndarray:
    """image_u8 [H,W,3] -> float32 [H,W] exactly as tta.py:196-213 builds ``preds``.
    ``blend="gaussian"`` restates the product's opt-in mode instead (numpy restatement only -- parity unpinned,
    there is no reference behaviour): preds = sum_t w_t * tile_t / sum_t w_t in float32, tiles in make_grid order."""
    H, W = image_u8.shape[:2]
    preds = np.zeros((H, W), dtype=np.float32)
    wsum = np.zeros((H, W), dtype=np.float32)
    g = gaussian_window(2 * S)
    w2d = (g[:, None] * g[None, :]).astype(np.float32)
    for (x1, x2, y1, y2) in make_grid((H, W), window=2 * S, min_overlap=32):
        tile = image_u8[x1:x2, y1:y2]
        tile = cv2.resize(tile, (S, S), interpolation=cv2.INTER_LINEAR)          # A.Resize(S, S)
        t = torch.from_numpy(preprocess(tile, mean, std).transpose(2, 0, 1)).float()[None]  # ToTensorV2 + .float()
        with torch.no_grad():
            logit = tta_mean_logits(net, t, tta_kind)[0][0]
            score = logit.sigmoid().cpu().numpy()
        score = cv2.resize(score, (2 * S, 2 * S), interpolation=cv2.INTER_LINEAR)
        if blend == "gaussian":
            preds[x1:x2, y1:y2] = preds[x1:x2, y1:y2] + w2d * score
            wsum[x1:x2, y1:y2] = wsum[x1:x2, y1:y2] + w2d
        else:
            preds[x1:x2, y1:y2] = score
    if blend == "gaussian":
        preds = np.where(wsum > 0, preds / np.where(wsum > 0, wsum, 1), 0).astype(np.float32)
    return preds


def whole_image_probability(prob_SxS: np.ndarray, crop_hw, ori_hw) -> np.ndarray:
    """tta.py:117-119: albumentations center_crop + cv2.resize(INTER_LINEAR) to the original size."""
    S_h, S_w = prob_SxS.shape
    ch, cw = crop_hw
    y0, x0 = (S_h - ch) // 2, (S_w - cw) // 2
    crop = prob_SxS[y0:y0 + ch, x0:x0 + cw]
    return cv2.resize(crop, (ori_hw[1], ori_hw[0]), interpolation=cv2.INTER_LINEAR)


def ensemble_probability(nets_list, x: torch.Tensor) -> np.ndarray:
    """x [1,3,S,S] -> float32 [S,S] as ensemble.py:86-100 forms ``mean_pred`` (D4 TTA on every model)."""
    mean_pred = None
    with torch.no_grad():
        for net in nets_list:
            pred = tta_mean_logits(net, x, "d4")
            pred = torch.sigmoid(pred)[0].squeeze(dim=0).numpy()
            if mean_pred is None:
                mean_pred = pred
            else:
                mean_pred += pred
    return mean_pred / len(nets_list)
