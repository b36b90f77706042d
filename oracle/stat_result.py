"""TEST INFRASTRUCTURE -- CPU restatement of src/main/stat_result.py:13-105 (export_result) and
stat_result_vessel.py, used only by tests/ as the checker of eyediseasesegmentation_b200.stat_result.

Pinned: tests/golden/stat_result.json holds the CSV files the REFERENCE'S OWN export_result wrote for a
seeded synthetic mask set (tests/golden/make_golden.py, loaded through oracle/ref_loader.py);
tests/test_oracle.py checks this restatement against them byte for byte.
"""
import os
import re

import numpy as np
from PIL import Image

EPS = 1e-7


def _binary(path):
    # stat_result.py:32-35 / 50-52: 'L' -> point(x > 50) in mode '1' -> uint8 {0, 1}
    im = Image.open(path).convert("L").point(lambda x: 255 if x > 50 else 0, "1")
    return np.asarray(im).astype(np.uint8)


def export_result(gt_dir, pred_dir, pred_name, save_dir):
    names = os.listdir(gt_dir)
    n = len(names)
    cols = {k: np.empty(n + 1, dtype=float) for k in ("sn", "ppv", "sp", "iou", "dice")}
    paths = np.empty(n + 1, dtype=object)
    for i, name in enumerate(names):
        paths[i] = name
        arr_gt = _binary(os.path.join(gt_dir, name))
        arr_pred = _binary(os.path.join(pred_dir, pred_name(name)))
        true_p = np.sum(arr_gt & arr_pred)
        actual_p = np.sum(arr_gt)
        pred_p = np.sum(arr_pred)
        false_p = pred_p - true_p
        actual_n = arr_gt.shape[0] * arr_gt.shape[1] - actual_p
        true_n = actual_n - false_p
        union = actual_p + false_p
        cols["sn"][i] = 1 if actual_p == 0 else float(true_p) / float(actual_p)
        cols["ppv"][i] = 1 if pred_p == 0 else float(true_p) / float(pred_p)
        cols["sp"][i] = 1 if actual_n == 0 else float(true_n) / float(actual_n)
        cols["iou"][i] = (true_p + EPS * (union == 0).astype("float")) / (actual_p + false_p + EPS)
        cols["dice"][i] = (2 * true_p + EPS * (union == 0).astype("float")) / (true_p + actual_p + false_p + EPS)
    paths[n] = "Avg:"
    for k in cols:
        cols[k][n] = np.mean(cols[k][:-1])
    os.makedirs(save_dir, exist_ok=True)
    for k, col in cols.items():
        np.savetxt(f"{save_dir}/{k}.csv", np.stack((paths, col), axis=1), delimiter=",", fmt="%s")


def lesion_pred_name(dataset_name, lesion_type):
    def f(image_path):
        if dataset_name == "IDRiD":
            return re.sub("_" + lesion_type + ".tif", ".jpg", image_path)
        return re.sub(".tif", ".jpg", image_path)
    return f
