"""Mirror of ``src/main/tta.py``: ``test_tta(logdir, config, args)`` (whole-image path, lines
56-148) and ``tta_patches(logdir, config, args)`` (sliding-window path, lines 150-238) with the
reference's config keys, checkpoint location, output directory and return behaviour, so
``pipeline.py`` can do ``from eyediseasesegmentation_b200.tta import *`` unchanged.

The arithmetic runs on the B200 kernels; decode (PIL) and the uint8 whole-image resize (cv2)
stay on the host as in the reference.  Differences that do not change results:
  - images are visited in sorted order (the reference shuffles its DataLoader, tta.py:84);
  - the three consumers share one inference pass (see ``_driver.CachedPredictions``);
  - with torchrun (one process per GPU) ``tta_patches`` deals (image, tile) units to the ranks and all-reduces
    the per-image integer histograms (``partition.py``); ``test_tta`` shards whole images and
    all-reduces the metric sums.
"""
from __future__ import annotations

import logging
import re
from pathlib import Path

import numpy as np
import torch

from . import archs, kernels as K, partition
from . import _driver as drv
from ._driver import get_model, str_2_bool, smp  # noqa: F401  (re-exported like the reference)
from .aucpr import get_auc, plot_aucpr_curve
from .util import lesion_dict, get_datapath, make_grid, multigen, save_output as so  # noqa: F401

__all__ = ["get_model", "str_2_bool", "test_tta", "tta_patches"]


def test_tta(logdir, config, args):
    import cv2
    img_paths, mask_paths = get_datapath(config["test_img_path"], config["test_mask_path"],
                                         lesion_type=config["lesion_type"])
    model = drv.build_model(config, logdir, args)
    preprocessing_fn, mean, std = archs.get_preprocessing_fn(dataset_name=config["dataset_name"],
                                                             grayscale=config["gray"])
    if config["gray"]:
        raise NotImplementedError("gray=True inputs are outside the B200 hot path (3-channel stem)")
    S = config["scale_size"]
    transforms = drv.tta_transforms(args)
    dev = drv.device()

    # TestSegmentation.__init__ (data/lesion_dataset.py:102-106): geometry from the first image
    first = drv.read_rgb(img_paths[0])
    ORI_H, ORI_W = first.shape[:2]
    CROP_H, CROP_W = drv.longest_max_size(first, 1024, cv2.INTER_LINEAR).shape[:2]
    pairs = drv.shard(list(zip(img_paths, mask_paths)))
    batch_size = config["val_batch_size"]

    def produce():
        for b0 in range(0, len(pairs), batch_size):
            chunk = pairs[b0:b0 + batch_size]
            x = torch.empty((len(chunk), 3, S, S), dtype=torch.float32).pin_memory()
            masks = []
            for i, (ip, mp) in enumerate(chunk):
                img = drv.pad_to_square(drv.longest_max_size(drv.read_rgb(ip), S, cv2.INTER_LINEAR), S)
                x[i] = torch.from_numpy(preprocessing_fn(img).transpose(2, 0, 1)).float()
                m = drv.read_mask(mp, 50)
                masks.append(drv.pad_to_square(drv.longest_max_size(m, S, cv2.INTER_NEAREST), S))
            prob = drv.predict_probs(model, transforms, x.to(dev, non_blocking=True))
            y0, x0 = (S - CROP_H) // 2, (S - CROP_W) // 2
            for i, (ip, _) in enumerate(chunk):
                full = torch.empty((ORI_H, ORI_W), dtype=torch.float32, device=dev)
                K.resize_paste(prob[i], full, (y0, x0, CROP_H, CROP_W), (0, 0), (ORI_H, ORI_W))
                crop_mask = masks[i][y0:y0 + CROP_H, x0:x0 + CROP_W]
                mask = cv2.resize(crop_mask, (ORI_W, ORI_H), interpolation=cv2.INTER_LINEAR)
                yield drv.scored(full, mask), mask, Path(ip).name

    predict_generator = drv.CachedPredictions(produce)

    logging.info("====> Estimate auc-pr score")
    mean_auc = get_auc(predict_generator(), config)
    logging.info(f"MEAN-AUC {mean_auc}")
    logging.info("====> Find optimal threshold from 0 to 1 w.r.t auc-pr curve")
    optim_thres1, optim_thres2, optim_thres3 = plot_aucpr_curve(predict_generator(), Path(logdir).name, config)
    logging.info(f"Optimal threshold is {optim_thres3}")
    logging.info("====> Output binary mask base on optimal threshold value")
    for pred_mask, _, mask_name in predict_generator():
        mask = (np.asarray(pred_mask) > optim_thres3).astype(np.uint8)
        so(mask, drv.output_dir(config, logdir) / mask_name)
    logging.info("====> Finishing inference")


def tta_patches(logdir, config, args):
    test_img_dir = config["test_img_path"]
    test_mask_dir = config["test_mask_path"] / lesion_dict[config["lesion_type"]].dir_name
    ALL_MASKS = sorted(test_mask_dir.glob("*.*"))
    model = drv.build_model(config, logdir, args)
    _, mean, std = archs.get_preprocessing_fn(dataset_name=config["dataset_name"], grayscale=config["gray"])
    if config["gray"]:
        # tta.py:166,196-204: this path always reads the RGB window; `gray` only swaps the per-channel statistics
        # for their luma-weighted scalars, broadcast over the three channels
        mean, std = [mean] * 3, [std] * 3
    transforms = drv.tta_transforms(args)
    resize_size = config["scale_size"]
    lesion = config["lesion_type"]

    def load(mask_path):
        img = test_img_dir / re.sub("_" + lesion + ".tif", ".jpg", mask_path.name)
        return drv.read_rgb(img), drv.read_mask(mask_path, 0)

    blend = drv.tile_blend_mode(config)
    rank, world_size = partition.world()
    if world_size > 1 and blend == "overwrite":
        # one process per GPU: (image, tile) units over the ranks, one all-reduce of the integer histograms
        produce = drv.partitioned_producer(model, transforms, ALL_MASKS, load, resize_size, mean, std)
    else:
        # one GPU -- or the opt-in Gaussian blend, whose overlaps need every tile of an image in one place: whole
        # images per rank, metric sums all-reduced by aucpr
        MY_MASKS = drv.shard(ALL_MASKS)

        def produce():
            # decode of image k+1 runs on a background thread while the GPU works on image k
            loaded = drv.prefetched([(lambda m=m: load(m)) for m in MY_MASKS])
            for mask_path, (image, gt_mask) in zip(MY_MASKS, loaded):
                pred, _ = drv.infer_image_host(model, transforms, torch.from_numpy(image), torch.from_numpy(gt_mask),
                                               resize_size, mean, std, blend=blend)
                yield pred, gt_mask, mask_path.name

    predict_generator = drv.CachedPredictions(produce)

    logging.info("====> Estimate auc-pr score")
    mean_auc = get_auc(predict_generator(), config)
    logging.info(f"MEAN-AUC {mean_auc}")
    logging.info("====> Find optimal threshold from 0 to 1 w.r.t auc-pr curve")
    optim_thres1, optim_thres2, optim_thres3 = plot_aucpr_curve(predict_generator(), Path(logdir).name, config)
    for mask_pred, _, mask_name in predict_generator():
        if np.asarray(mask_pred).size == 0:         # partitioned run: another rank assembles and writes this image
            continue
        mask = (np.asarray(mask_pred) > optim_thres3).astype(np.float32)
        mask_name = re.sub("_" + config["lesion_type"] + ".tif", ".jpg", mask_name)
        so(mask, drv.output_dir(config, logdir) / mask_name)
    logging.info("====> Finishing inference")
