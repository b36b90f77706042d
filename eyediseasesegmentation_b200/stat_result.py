"""Mirror of ``src/main/stat_result.py`` (``export_result(save_dir, test_config)``, lines 13-105) and of
its vessel twin ``src/main/stat_result_vessel.py``: per-image SN / PPV / SP / IoU / Dice of the saved
binary masks against the ground truth, written as the same five CSV files.

``pipeline.py:107`` calls it right after the inference-and-scoring hot path.
The reference binarises both images with ``x > 50`` and counts with numpy on the host; here the two
uint8 maps go to the GPU once and ``eds_confusion_u8`` returns the three integers every metric is made
of (true_p, actual_p, pred_p).  File discovery, naming rules, metric formulas, the "Avg:" row and the
CSV formatting are the reference's, so the files are byte-identical.
"""
from __future__ import annotations

import os
import re

import numpy as np
import torch

from . import kernels as K
from .util import lesion_dict

EPS = 1e-7


def _read_l(path) -> np.ndarray:
    from PIL import Image
    return np.array(Image.open(path).convert("L"), dtype=np.uint8)


def confusion_on_device(pairs):
    """[(pred u8 HxW, gt u8 HxW), ...] -> int64 [n, 3] (true_p, actual_p, pred_p), masks = x > 50.
    One launch per image size group; a single device->host read at the end."""
    if not torch.cuda.is_available():
        raise RuntimeError("export_result counts on the GPU (eds_confusion_u8); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.zeros((len(pairs), 3), dtype=torch.int64, device=dev)
    for i, (pred, gt) in enumerate(pairs):
        if pred.shape != gt.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {gt.shape} {pred.shape}")
        p = torch.from_numpy(np.ascontiguousarray(pred, dtype=np.uint8)).to(dev, non_blocking=True).view(1, -1)
        g = torch.from_numpy(np.ascontiguousarray(gt, dtype=np.uint8)).to(dev, non_blocking=True).view(1, -1)
        K.confusion_counts(p, g, 50, 50, counts=out[i:i + 1])
    return out.cpu().numpy()


def metrics_from_counts(true_p, actual_p, pred_p, n_pixels):
    """stat_result.py:55-79, same operations in the same order on numpy integer scalars."""
    true_p, actual_p, pred_p = np.int64(true_p), np.int64(actual_p), np.int64(pred_p)
    false_p = pred_p - true_p
    actual_n = n_pixels - actual_p
    true_n = actual_n - false_p
    union = actual_p + false_p
    sn = 1 if actual_p == 0 else float(true_p) / float(actual_p)
    ppv = 1 if pred_p == 0 else float(true_p) / float(pred_p)
    sp = 1 if actual_n == 0 else float(true_n) / float(actual_n)
    iou = (true_p + EPS * (union == 0).astype("float")) / (actual_p + false_p + EPS)
    dice = (2 * true_p + EPS * (union == 0).astype("float")) / (true_p + actual_p + false_p + EPS)
    return sn, ppv, sp, iou, dice


def _export(gt_dir: str, pred_dir: str, pred_name, save_dir: str):
    names = os.listdir(gt_dir)
    test_size = len(names)
    sn = np.empty(test_size + 1, dtype=float)
    ppv = np.empty(test_size + 1, dtype=float)
    sp = np.empty(test_size + 1, dtype=float)
    dice = np.empty(test_size + 1, dtype=float)
    iou = np.empty(test_size + 1, dtype=float)
    image_paths = np.empty(test_size + 1, dtype=object)

    pairs = []
    for image_path in names:
        gt = _read_l(gt_dir + "/" + image_path)
        pred = _read_l(pred_dir + "/" + pred_name(image_path))
        pairs.append((pred, gt))
    counts = confusion_on_device(pairs) if pairs else np.zeros((0, 3), dtype=np.int64)
    for i, image_path in enumerate(names):
        image_paths[i] = image_path
        gt = pairs[i][1]
        sn[i], ppv[i], sp[i], iou[i], dice[i] = metrics_from_counts(counts[i, 0], counts[i, 1], counts[i, 2],
                                                                     gt.shape[0] * gt.shape[1])
    i = test_size
    image_paths[i] = "Avg:"
    sn[i] = np.mean(sn[:-1])
    ppv[i] = np.mean(ppv[:-1])
    sp[i] = np.mean(sp[:-1])
    iou[i] = np.mean(iou[:-1])
    dice[i] = np.mean(dice[:-1])

    if not os.path.exists(save_dir):
        os.makedirs(save_dir, exist_ok=True)
    for name, col in (("sn", sn), ("ppv", ppv), ("sp", sp), ("iou", iou), ("dice", dice)):
        np.savetxt(f"{save_dir}/{name}.csv", np.stack((image_paths, col), axis=1), delimiter=",", fmt="%s")
    print(f"Results are saved at {save_dir}")


def export_result(save_dir, test_config):
    """stat_result.py:13-105."""
    gt_dir = str(test_config["test_mask_path"] / lesion_dict[test_config["lesion_type"]].dir_name)
    pred_dir = test_config["out_dir"] + "/" + test_config["dataset_name"] + "/tta/" + save_dir

    def pred_name(image_path):
        if test_config["dataset_name"] == "IDRiD":
            return re.sub("_" + test_config["lesion_type"] + ".tif", ".jpg", image_path)
        return re.sub(".tif", ".jpg", image_path)

    out = test_config["out_dir"] + "/" + test_config["dataset_name"] + "/result_assessment/" + save_dir
    _export(gt_dir, pred_dir, pred_name, out)


def export_result_vessel(save_dir, test_config):
    """stat_result_vessel.py: masks directly under test_mask_path, predictions saved under the
    mask's own file name (:16, :43)."""
    gt_dir = str(test_config["test_mask_path"])
    pred_dir = test_config["out_dir"] + "/" + test_config["dataset_name"] + "/tta/" + save_dir
    out = test_config["out_dir"] + "/" + test_config["dataset_name"] + "/result_assessment/" + save_dir
    _export(gt_dir, pred_dir, lambda image_path: image_path, out)
