"""Mirror of the reference's top-level ``ensemble.py``: ``get_model``, ``get_best_model(path)``
and ``predict(config, logdirs, outdir)`` -- several trained models, each under D4 test-time augmentation, their
per-image probability maps averaged, then AUC-PR scoring and mask export.

Arithmetic (ensemble.py:86-100): for every image (batch 1, LongestMaxSize + centred pad to S x S),
``mean_pred = sum over models, in list order, of sigmoid(D4-mean logits)`` divided by the model count, in fp32,
kept at S x S (this driver does not resize back).  On the B200 each model runs its fused TTA forward
(stem loader folds the views, ``eds_tta_merge`` de-augments + means + sigmoids) and the model axis is reduced by
the same merge kernel with identity view maps (same running-sum order, so the mean is bit-identical to the
reference's ``+=`` loop on identical per-model maps).

Where the reference file is inconsistent with the rest of the reference it is mirrored by intent, not by letter:
  - it calls ``get_auc(gt_masks, tta_predictions, config)`` / ``plot_aucpr_curve(gt_masks, tta_predictions,
    outdir, config)`` (ensemble.py:102,105-109) although ``aucpr.py:17,45`` take a ``(pred, gt, name)`` generator
    and return three thresholds; here the generator form is used and the first threshold is applied, which is
    what ``pred_mask > optim_thres1`` (ensemble.py:114) names;
  - ``TestSegmentation(img_paths, mask_paths, transform=...)`` (ensemble.py:78) passes the masks in the
    ``is_gray`` slot; here masks are masks;
  - ``archs.get_preprocessing_fn(dataset_name=...)`` (ensemble.py:73) omits the required ``grayscale`` argument
    (archs/__init__.py:61); RGB is what the rest of that file assumes;
  - the dataset location is hard-coded (ensemble.py:65-66); ``config['test_img_path']`` /
    ``config['test_mask_path']`` override it when present.
"""
from __future__ import annotations

import json
import logging
import os
from pathlib import Path

import numpy as np
import torch

from . import archs, kernels as K, ttach_compat as tta
from . import _driver as drv
from ._driver import get_model, smp  # noqa: F401  (re-exported like the reference)
from .aucpr import get_auc, plot_aucpr_curve
from .util import get_datapath, save_output as so

__all__ = ["get_model", "get_best_model", "predict", "ensemble_mean"]

MAX_MODELS = 8      # eds_tta_merge reduces up to 8 planes per launch


def get_best_model(path):
    """ensemble.py:39-62: ``path/config.json`` names the architecture, ``path/checkpoints/best.pth`` holds the
    weights; the model comes back in eval mode on the GPU, wrapped for D4 TTA with mean merging."""
    path = Path(path)
    checkpoint = drv.load_checkpoint(path / "checkpoints/best.pth")
    with open(path / "config.json", "r") as j:
        config = json.load(j)
    if hasattr(smp, config["model_name"]):
        model = get_model(config["model_params"], config["model_name"])
    elif config["model_name"] == "TransUnet":
        raise NotImplementedError("TransUnet is outside the B200 hot path")
    else:
        model = archs.get_model(model_name=config["model_name"], params=config["model_params"], training=False)
    model.load_state_dict(checkpoint["model_state_dict"])
    model.eval()
    model = model.to(drv.device())
    return tta.SegmentationTTAWrapper(model, tta.aliases.d4_transform(), merge_mode="mean")


def ensemble_mean(probs: torch.Tensor) -> torch.Tensor:
    """probs [M,B,S,S] fp32 on the device -> [B,S,S]: running sum in model order / M (ensemble.py:94-96)."""
    M = probs.shape[0]
    if M > MAX_MODELS:
        raise ValueError(f"ensemble of {M} models: at most {MAX_MODELS} are reduced per launch")
    identity = [(1, 0, 0, 0, 1, 0)] * M
    return K.tta_merge(probs.contiguous(), identity, False)


def predict(config, logdirs, outdir):
    import cv2
    test_img_dir = Path(config.get("test_img_path", "data/raw/IDRiD/1. Original Images/b. Testing Set"))
    test_mask_dir = Path(config.get("test_mask_path", "data/raw/IDRiD/2. All Segmentation Groundtruths/b. Testing Set"))
    img_paths, mask_paths = get_datapath(test_img_dir, test_mask_dir, lesion_type=config["lesion_type"])

    models = [get_best_model(logdir) for logdir in logdirs]
    preprocessing_fn, _, _ = archs.get_preprocessing_fn(dataset_name=config["dataset_name"], grayscale=False)
    S = config["scale_size"]
    dev = drv.device()
    pairs = drv.shard(list(zip(img_paths, mask_paths)))

    def produce():
        for ip, mp in pairs:
            img = drv.pad_to_square(drv.longest_max_size(drv.read_rgb(ip), S, cv2.INTER_LINEAR), S)
            x = drv._pinned((1, 3, S, S), torch.float32, "ens_x")
            x[0] = torch.from_numpy(preprocessing_fn(img).transpose(2, 0, 1)).float()
            mask = drv.pad_to_square(drv.longest_max_size(drv.read_mask(mp, 50), S, cv2.INTER_NEAREST), S)
            xd = x.to(dev, non_blocking=True)
            # each wrapper returns the D4-mean LOGITS (ensemble.py:91); the sigmoid is applied per model (:93)
            probs = torch.stack([drv.predict_probs(m.model, m.transforms, xd) for m in models])
            mean_pred = ensemble_mean(probs)[0]
            yield drv.scored(mean_pred, mask), mask, Path(ip).name

    predictions = drv.CachedPredictions(produce)

    logging.info("====> Estimate auc-pr score")
    mean_auc = get_auc(predictions(), config)
    logging.info(f"MEAN-AUC {mean_auc}")
    logging.info("====> Find optimal threshold from 0 to 1 w.r.t auc-pr curve")
    optim_thres1, optim_thres2, _ = plot_aucpr_curve(predictions(), outdir, config)
    logging.info(f"Optimal threshold is {optim_thres1}")
    logging.info("====> Output binary mask base on optimal threshold value")
    out_path = Path(config["out_dir"]) / config["dataset_name"] / "tta" / config["lesion_type"] / outdir
    if not os.path.isdir(out_path):
        os.makedirs(out_path, exist_ok=True)
    for pred_mask, _, mask_name in predictions():
        mask = (np.asarray(pred_mask) > optim_thres1).astype(np.uint8)
        so(mask, out_path / mask_name)
    logging.info("====> Finishing inference")
    return mean_auc
