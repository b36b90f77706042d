"""Scoring interface of the reference (``src/main/aucpr.py``) on the B200 histogram kernels.

Same four entry points, same arguments, same return values:

  get_auc(generator, config) -> float                    aucpr.py:17-30
  get_aucroc(generator, config) -> float                 aucpr.py:32-43
  plot_aucpr_curve(generator, exp_name, test_config)     aucpr.py:45-118  -> (t_absdiff, t_dist, t_fscore)
  plot_aucroc_curve(generator, exp_name, test_config)    aucpr.py:120-205 -> t

``generator`` is any iterable of ``(pred float32[H,W], gt uint8[H,W] in {0,1}, name)``.
Where the reference sorts 12 M scores per image with sklearn and makes 19 numpy passes per
image, this module bins every pixel once on the GPU (``eds_pr_hist_f32``) and reads AP,
ROC-AUC and all 19 threshold counts from a scan of the bins (``eds_pr_scan``).  Threshold
counts are exact integers; AP / ROC-AUC are exact on scores quantised to the histogram key (symmetric about 1/2, 9 mantissa bits per binade of min(p, 1-p)) and
within 1e-3 (measured ~1e-4) of sklearn on the raw fp32 scores.

Arrays yielded by this package's own drivers carry their scores already (computed while the
probability map was still in HBM), so the reference's three passes over the test set cost
one inference pass and one histogram pass.
"""
from __future__ import annotations

import logging
import os
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib, kernels as K

logging.basicConfig(level=logging.INFO)

thresh_list = list(_lib.PR_THRESHOLDS)


class ImageScores:
    """Per-image result of the histogram + scan kernels."""
    __slots__ = ("ap", "roc", "tp", "pp", "n_pos", "n_neg", "replicated")

    def __init__(self, ap, roc, tp, pp, n_pos, n_neg, replicated=False):
        self.ap, self.roc, self.tp, self.pp, self.n_pos, self.n_neg = ap, roc, tp, pp, n_pos, n_neg
        #: True when every rank already holds this image's GLOBAL scores (tile-partitioned path: the integer
        #: histograms were all-reduced before the scan), so the entry points must not sum over ranks again
        self.replicated = replicated


class ScoredArray(np.ndarray):
    """float32 probability map that remembers the scores computed on the device."""

    def __new__(cls, arr, scores: Optional[ImageScores] = None):
        obj = np.asarray(arr).view(cls)
        obj._eds_scores = scores
        return obj

    def __array_finalize__(self, obj):
        self._eds_scores = None  # derived arrays (comparisons, slices) are not scored


def score_device(prob: torch.Tensor, gt: torch.Tensor) -> ImageScores:
    """prob [H,W] fp32 cuda, gt [H,W] uint8 cuda -> ImageScores (one small D2H read)."""
    hist, strad = K.pr_hist(prob.reshape(1, -1), gt.reshape(1, -1))
    ap, roc, counts, totals = K.pr_scan(hist, strad)
    counts = counts.cpu().numpy()
    totals = totals.cpu().numpy()
    return ImageScores(float(ap.item()), float(roc.item()), counts[0, :, 0].copy(), counts[0, :, 1].copy(),
                       int(totals[0, 0]), int(totals[0, 1]))


def _scores(pred, gt) -> ImageScores:
    cached = getattr(pred, "_eds_scores", None)
    if cached is not None:
        return cached
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    gt = np.ascontiguousarray(gt)
    if pred.shape != gt.shape:
        raise ValueError(f"pred {pred.shape} and gt {gt.shape} differ in shape")
    if gt.dtype != np.uint8:
        gt = gt.astype(np.uint8)
    if gt.size and gt.max() > 1:
        raise ValueError("gt mask must be binary {0,1} (tta.py:192-194 binarises it)")
    dev = torch.device("cuda", torch.cuda.current_device())
    return score_device(torch.from_numpy(pred).to(dev, non_blocking=True), torch.from_numpy(gt).to(dev, non_blocking=True))


def _dist_sum(values, dtype, replicated=False):
    """Sum over ranks when torch.distributed is initialised (one process per GPU, images
    sharded by the drivers).  This is the path's single collective: a few integers / two
    float64 -- NCCL over NVLink on the GPU box, gloo in the CPU tests."""
    import torch.distributed as dist
    if replicated or not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return values
    on_gpu = dist.get_backend() == "nccl"
    t = torch.tensor(values, dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()) if on_gpu else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def _all_replicated(flags) -> bool:
    """Every item carried global scores (tile-partitioned drivers); a mix would double count."""
    if any(flags) and not all(flags):
        raise ValueError("items with replicated (already all-reduced) and rank-local scores cannot be mixed")
    return bool(flags) and all(flags)


def get_auc(generator: Iterable, config):
    sum_pav = 0
    i = 0
    replicated = []
    for pred_mask, gt_mask, _ in generator:
        s = _scores(pred_mask, gt_mask)
        replicated.append(s.replicated)
        if s.n_pos == 0:
            continue
        pav = s.ap
        print("PAV", pav)
        sum_pav += pav
        i += 1
    sum_pav, i = _dist_sum([float(sum_pav), float(i)], torch.float64, _all_replicated(replicated))
    mpav = sum_pav / i   # ZeroDivisionError when no image has positives, as in the reference
    return mpav


def get_aucroc(generator: Iterable, config):
    sum_pav = 0
    i = 0
    one_class = 0
    replicated = []
    for pred_mask, gt_mask, _ in generator:
        s = _scores(pred_mask, gt_mask)
        replicated.append(s.replicated)
        if s.n_pos == 0:
            continue
        if s.n_neg == 0:
            one_class += 1      # raised AFTER the collective: an exception on one rank must not strand the others
            continue
        sum_pav += s.roc
        i += 1
    sum_pav, i, one_class = _dist_sum([float(sum_pav), float(i), float(one_class)], torch.float64,
                                      _all_replicated(replicated))
    if one_class:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return sum_pav / i


def _trapezoid_auc(x: np.ndarray, y: np.ndarray) -> float:
    """sklearn.metrics.auc: trapezoid rule, x must be monotonic (either direction)."""
    if x.shape[0] < 2:
        raise ValueError("At least 2 points are needed to compute area under curve")
    dx = np.diff(x)
    direction = 1
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1
        else:
            raise ValueError("x is neither increasing nor decreasing : {}.".format(x))
    return float(direction * np.sum((y[1:] + y[:-1]) * 0.5 * dx))


def _first_sorted(values, reverse):
    """``sorted(zip(values, thresh_list), key=first, reverse=...)[0][1]`` of aucpr.py:93-98."""
    return sorted(list(zip(values, thresh_list)), key=lambda t: t[0], reverse=reverse)[0][1]


def _pooled_counts(generator):
    tp = np.zeros(len(thresh_list), dtype=np.int64)
    pp = np.zeros(len(thresh_list), dtype=np.int64)
    ap = 0
    an = 0
    n_images = 0
    replicated = []
    for pred_mask, gt_mask, _ in generator:
        s = _scores(pred_mask, gt_mask)
        replicated.append(s.replicated)
        tp += s.tp.astype(np.int64)
        pp += s.pp.astype(np.int64)
        ap += s.n_pos
        an += s.n_neg
        n_images += 1
    flat = _dist_sum(tp.tolist() + pp.tolist() + [int(ap), int(an), n_images], torch.int64,
                     _all_replicated(replicated))
    n = len(thresh_list)
    tp, pp = np.array(flat[:n], dtype=np.int64), np.array(flat[n:2 * n], dtype=np.int64)
    ap, an, n_images = flat[2 * n], flat[2 * n + 1], flat[2 * n + 2]
    if n_images == 0:
        raise KeyError(str(np.array(thresh_list)[0]))  # the reference indexes an empty dict here
    return tp, pp, ap, an


def _figure(figure_dir, exp_name, x, y, title, labels):
    """The plot is a side effect of the reference (plotly + orca); it is written when plotly is
    importable and skipped otherwise -- the returned thresholds do not depend on it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
        return                  # every rank holds the same pooled curve; one writer
    try:
        import plotly.express as px
    except Exception:
        logging.info("plotly is not installed: skipping the curve image")
        return
    fig = px.area(x=x, y=y, title=title, labels=labels, width=700, height=500)
    fig.add_shape(type="line", line=dict(dash="dash"), x0=0, x1=1, y0=1, y1=0)
    fig.update_yaxes(scaleanchor="x", scaleratio=1)
    fig.update_xaxes(constrain="domain")
    fig.write_image(figure_dir + "/{}.jpg".format(str(exp_name)))


def pr_curve_from_counts(tp, pp, ap):
    """Rates of aucpr.py:83-91 from pooled integer counts."""
    sn = (tp.astype(float) + 1e-7) / (float(ap) + 1e-7)
    ppv = (tp.astype(float) + 1e-7) / (pp.astype(float) + 1e-7)
    f_score = (2 * sn * ppv) / (sn + ppv)
    return sn, ppv, f_score, _trapezoid_auc(sn, ppv)


def plot_aucpr_curve(generator: Iterable, exp_name, test_config):
    figure_dir = os.path.join(test_config["out_dir"], test_config["dataset_name"], "figures", test_config["lesion_type"])
    if not os.path.exists(figure_dir):
        os.makedirs(figure_dir, exist_ok=True)   # exist_ok: several ranks may arrive together
    tp, pp, ap, _ = _pooled_counts(generator)
    recall, precision, f_score, aucpr = pr_curve_from_counts(tp, pp, ap)
    optimal_threshold = _first_sorted(np.abs(precision - recall), reverse=False)
    optimal_threshold_1 = _first_sorted(np.sqrt((1 - precision) ** 2 + (1 - recall) ** 2), reverse=False)
    optimal_threshold_2 = _first_sorted(f_score, reverse=True)
    logging.info(f"OPTIMAL THRESHOLD: {optimal_threshold}")
    logging.info(f"OPTIMAL THRESHOLD 1: {optimal_threshold_1}")
    logging.info(f"OPTIMAL THRESHOLD 2: {optimal_threshold_2}")
    _figure(figure_dir, exp_name, recall, precision,
            f"Precision-Recall Curve AUC:{aucpr}-Optimal threshold: {optimal_threshold_2}",
            dict(x="Recall", y="Precision"))
    logging.info(f"Saved AUC-PR Curve to {figure_dir}")
    return optimal_threshold, optimal_threshold_1, optimal_threshold_2


def roc_curve_from_counts(tp, pp, ap, an):
    """Rates of aucpr.py:173-184: tn = an - (pp - tp)."""
    tn = an - (pp - tp)
    sn = (tp.astype(float) + 1e-7) / (float(ap) + 1e-7)
    sp = (tn.astype(float) + 1e-7) / (float(an) + 1e-7)
    ppv = (tp.astype(float) + 1e-7) / (pp.astype(float) + 1e-7)
    tpr, fpr = sn, 1 - sp
    f_score = (2 * tpr * ppv) / (tpr + ppv)
    return tpr, fpr, f_score, _trapezoid_auc(fpr, tpr)


def plot_aucroc_curve(generator: Iterable, exp_name, test_config):
    figure_dir = os.path.join(test_config["out_dir"], test_config["dataset_name"], "figures")
    if not os.path.exists(figure_dir):
        os.makedirs(figure_dir, exist_ok=True)   # exist_ok: several ranks may arrive together
    tp, pp, ap, an = _pooled_counts(generator)
    tpr, fpr, f_score, aucroc = roc_curve_from_counts(tp, pp, ap, an)
    optimal_threshold = thresh_list[int(np.argmax(f_score))]
    logging.info(f"OPTIMAL THRESHOLD: {optimal_threshold}")
    _figure(figure_dir, exp_name, fpr, tpr, f"ROC Curve AUC:{aucroc}-Optimal threshold: {optimal_threshold}",
            dict(x="False Positive Rate", y="True Positive Rate"))
    logging.info(f"Saved AUC-ROC Curve to {figure_dir}")
    return optimal_threshold
