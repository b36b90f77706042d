"""Mirror of ``src/main/util/aucpr_cb.py``: ``AucPRMetricCallback``, the validation-epoch
PR-AUC of the training loop, with catalyst's callback protocol (``on_loader_start`` / ``on_batch_end`` /
``on_loader_end`` on a ``runner`` that carries ``output``, ``input`` and ``loader_metrics``).

The reference keeps every probability map and label of the loader on the host, pickles them through
``all_gather`` and sorts the lot in ``sklearn.precision_recall_curve`` (aucpr_cb.py:52-64).  Here each batch is
binned on the GPU into ONE pooled score histogram (``eds_pr_hist_f32``, the kernel of the test-time AUC-PR) while
the activations are still in HBM; ranks exchange 2 x 23 556 integers with one all-reduce, and the curve area is
read off the histogram:

    thresholds = occupied bins in descending score order,  tps / fps = running sums,
    precision = tps / (tps + fps),  recall = tps / tps[-1],  curve prepended with (recall 0, precision 1),
    score = trapezoid area  (``sklearn.metrics.auc(recall, precision)``, aucpr_cb.py:63-64).

Scores sharing a bin (min(p, 1-p) equal in its top 9 mantissa bits) are treated as tied, which moves the area by < 1e-4
on the test sets (tests/test_host_logic.py, tests/test_golden_gpu.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor

__all__ = ["AucPRMetricCallback", "pr_auc_from_hist"]

try:                                              # the training loop's framework; absent in the B200 image
    from catalyst.core import Callback as _Callback, CallbackOrder as _Order
    _ORDER = _Order.Metric
except Exception:                                 # same protocol without the dependency
    class _Callback:                              # noqa: D401
        def __init__(self, order=None):
            self.order = order
    _ORDER = 80                                   # catalyst's CallbackOrder.Metric


def pr_auc_from_hist(neg: np.ndarray, pos: np.ndarray) -> float:
    """Area under the precision-recall curve from per-bin counts (ascending score order), following
    ``precision_recall_curve`` + ``auc`` of scikit-learn 0.24 as called at aucpr_cb.py:63-64."""
    pos = np.asarray(pos, dtype=np.int64)[::-1]
    neg = np.asarray(neg, dtype=np.int64)[::-1]
    occupied = (pos + neg) > 0
    tps = np.cumsum(pos[occupied]).astype(np.float64)
    fps = np.cumsum(neg[occupied]).astype(np.float64)
    if tps.size == 0 or tps[-1] == 0:
        return float("nan")                       # sklearn: recall = tps / 0
    precision = np.concatenate([[1.0], tps / (tps + fps)])
    recall = np.concatenate([[0.0], tps / tps[-1]])
    return float(np.sum(np.diff(recall) * (precision[1:] + precision[:-1]) * 0.5))


class AucPRMetricCallback(_Callback):
    """Auc Precision-Recall score metric (same constructor as the reference, aucpr_cb.py:20-46)."""

    def __init__(
        self,
        outputs_to_probas: Callable[[Tensor], Tensor] = torch.sigmoid,
        input_key: str = "targets",
        output_key: str = "logits",
        prefix: str = "auc_pr",
        average="macro",
        ignore_index: Optional[int] = None,
    ):
        super().__init__(_ORDER)
        self.prefix = prefix
        self.output_key = output_key
        self.input_key = input_key
        self.ignore_index = ignore_index
        self.outputs_to_probas = outputs_to_probas
        self.average = average
        self._hist = None
        self._straddle = None

    def on_loader_start(self, state):
        self._hist = None
        self._straddle = None

    @torch.no_grad()
    def on_batch_end(self, runner):
        from . import kernels as K
        pred_probas = self.outputs_to_probas(runner.output[self.output_key])
        true_labels = runner.input[self.input_key]
        if not pred_probas.is_cuda:
            raise RuntimeError("AucPRMetricCallback bins on the GPU; there is no CPU fallback")
        prob = pred_probas.detach().reshape(1, -1).to(torch.float32).contiguous()
        # precision_recall_curve: positives are the entries equal to pos_label = 1
        gt = (true_labels.to(prob.device).reshape(1, -1) == 1).to(torch.uint8).contiguous()
        # the kernel's counters are u32 stored in an int32 tensor: fold every batch into an int64 accumulator
        # (mask off the sign extension) so a bin can grow past 2^31 over a long loader
        hist, _ = K.pr_hist(prob, gt)
        wide = hist.to(torch.int64) & 0xFFFFFFFF
        self._hist = wide if self._hist is None else self._hist + wide

    def on_loader_end(self, runner):
        import torch.distributed as dist
        if self._hist is None:
            raise ValueError("need at least one array to concatenate")     # np.concatenate([]) in the reference
        hist = self._hist.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(hist)
        h = hist.cpu().numpy()[0]
        runner.loader_metrics[self.prefix] = pr_auc_from_hist(h[0], h[1])
