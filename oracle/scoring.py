"""TEST INFRASTRUCTURE: CPU restatement of the reference's scoring (src/main/aucpr.py).

sklearn (installed here and on the GPU box; the reference pins 0.24.1, this image has 1.9)
provides ``average_precision_score`` / ``roc_auc_score`` exactly as the reference calls them
(aucpr.py:24,38); the 19-threshold loops (aucpr.py:60-81,136-170) are restated in numpy.
tests/test_oracle.py checks every function here against the reference's own aucpr.py
(loaded through oracle/ref_loader.py) and tests/golden/ holds the reference's outputs.
"""
from __future__ import annotations

import numpy as np
from sklearn.metrics import auc, average_precision_score, roc_auc_score

THRESH_LIST = [0, 0.00001, 0.0001, 0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.99, 0.999,
               0.9999, 0.99999, 1]


def get_auc(items) -> float:
    """aucpr.py:17-30: mean AP over the images that contain positives."""
    aps = [average_precision_score(gt.reshape(-1), pred.reshape(-1)) for pred, gt, _ in items if gt.sum() != 0]
    return sum(aps) / len(aps)


def get_aucroc(items) -> float:
    """aucpr.py:32-43."""
    vals = [roc_auc_score(gt.reshape(-1), pred.reshape(-1)) for pred, gt, _ in items if gt.sum() != 0]
    return sum(vals) / len(vals)


def threshold_counts(pred: np.ndarray, gt: np.ndarray):
    """tp / actual-p / pred-p per threshold for one image (aucpr.py:61-66; strict '>',
    float32 score against float64 threshold)."""
    th = np.array(THRESH_LIST)
    tp = np.zeros(len(th), dtype=np.int64)
    pp = np.zeros(len(th), dtype=np.int64)
    for k in range(len(th)):
        above = (pred > th[k]).astype("uint8")
        tp[k] = np.sum(gt & above)
        pp[k] = np.sum(above)
    return tp, int(np.sum(gt)), pp


def pooled_counts(items):
    tp = np.zeros(len(THRESH_LIST), dtype=np.int64)
    pp = np.zeros(len(THRESH_LIST), dtype=np.int64)
    ap = 0
    an = 0
    for pred, gt, _ in items:
        t, a, p = threshold_counts(pred, gt)
        tp += t
        pp += p
        ap += a
        an += gt.shape[0] * gt.shape[1] - a
    return tp, pp, ap, an


def pr_curve(items):
    """aucpr.py:83-98 -> dict(recall, precision, aucpr, thresholds=(absdiff, dist, fscore))."""
    tp, pp, ap, _ = pooled_counts(items)
    recall = (tp.astype(float) + 1e-7) / (float(ap) + 1e-7)
    precision = (tp.astype(float) + 1e-7) / (pp.astype(float) + 1e-7)
    f_score = (2 * recall * precision) / (recall + precision)
    first = lambda vals, rev: sorted(list(zip(vals, THRESH_LIST)), key=lambda i: i[0], reverse=rev)[0][1]  # noqa: E731
    return dict(tp=tp, pp=pp, ap=ap, recall=recall, precision=precision, aucpr=auc(recall, precision),
                thresholds=(first(np.abs(precision - recall), False),
                            first(np.sqrt((1 - precision) ** 2 + (1 - recall) ** 2), False),
                            first(f_score, True)))


def roc_curve(items):
    """aucpr.py:173-186 -> dict(tpr, fpr, aucroc, threshold)."""
    tp, pp, ap, an = pooled_counts(items)
    tn = an - (pp - tp)
    tpr = (tp.astype(float) + 1e-7) / (float(ap) + 1e-7)
    sp = (tn.astype(float) + 1e-7) / (float(an) + 1e-7)
    precision = (tp.astype(float) + 1e-7) / (pp.astype(float) + 1e-7)
    fpr = 1 - sp
    f_score = (2 * tpr * precision) / (tpr + precision)
    return dict(tp=tp, pp=pp, ap=ap, an=an, tpr=tpr, fpr=fpr, aucroc=auc(fpr, tpr),
                threshold=THRESH_LIST[int(np.argmax(f_score))])


def score_key(pred: np.ndarray) -> np.ndarray:
    """The product's histogram key (include/eds_b200.h, EDS_PR_*) restated in numpy, so tests
    can ask sklearn for the AP of key-quantised scores."""
    p = np.ascontiguousarray(pred, dtype=np.float32)
    hi = p >= np.float32(0.5)
    q = np.where(hi, np.float32(1.0) - p, p).astype(np.float32)
    half = 23 * 512 + 2
    k = np.clip((q.view(np.int32).astype(np.int64) >> 14) - ((103 << 9) - 1), 0, half - 1)
    return np.where(hi, 2 * half - 1 - k, k)


def callback_pr_auc(y_trues, y_preds) -> float:
    """src/main/util/aucpr_cb.py:59-65: trapezoid area under sklearn's precision_recall_curve over every
    pixel of the loader (lists of per-sample arrays, concatenated)."""
    from sklearn.metrics import auc, precision_recall_curve
    y_trues = np.concatenate([np.asarray(t) for t in y_trues])
    y_preds = np.concatenate([np.asarray(p) for p in y_preds])
    precision, recall, _ = precision_recall_curve(y_trues.reshape(-1), y_preds.reshape(-1))
    return float(auc(recall, precision))
