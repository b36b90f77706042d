"""Mirror of ``src/main/tta_vessel.py`` (vessel pipeline: DRIVE / CHASEDB1 / HRF):
``test_tta`` (lines 55-136: whole image, batch 1, no resize -- inputs are pre-padded squares,
default IDRiD mean/std because ``dataset_name=None`` is passed at line 73) and
``tta_patches`` (lines 138-229), both scored with the ROC twins ``get_aucroc`` /
``plot_aucroc_curve``.
"""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from . import archs, partition
from . import _driver as drv
from ._driver import get_model, str_2_bool, smp  # noqa: F401
from .aucpr import get_aucroc, plot_aucroc_curve
from .util import lesion_dict, get_datapath, make_grid, multigen, save_output as so  # noqa: F401

__all__ = ["get_model", "str_2_bool", "test_tta", "tta_patches"]


def _finish(predict_generator, logdir, config, as_float):
    logging.info("====> Estimate auc-roc score")
    mean_auc = get_aucroc(predict_generator(), config)
    logging.info(f"MEAN-AUC {mean_auc}")
    logging.info("====> Find optimal threshold from 0 to 1 w.r.t auc-roc curve")
    optim_thres = plot_aucroc_curve(predict_generator(), Path(logdir).name, config)
    logging.info(f"Optimal threshold is {optim_thres}")
    for pred_mask, _, mask_name in predict_generator():
        if np.asarray(pred_mask).size == 0:           # partitioned run: another rank assembles and writes this image
            continue
        mask = (np.asarray(pred_mask) > optim_thres).astype(np.float32 if as_float else np.uint8)
        so(mask, drv.output_dir(config, logdir) / mask_name)
    logging.info("====> Finishing inference")


def test_tta(logdir, config, args):
    img_paths, mask_paths = get_datapath(config["test_img_path"], config["test_mask_path"],
                                         lesion_type=config["lesion_type"])
    model = drv.build_model(config, logdir, args)
    preprocessing_fn, mean, std = archs.get_preprocessing_fn(dataset_name=None, grayscale=config["gray"])
    if config["gray"]:
        raise NotImplementedError("gray=True inputs are outside the B200 hot path (3-channel stem)")
    transforms = drv.tta_transforms(args)
    dev = drv.device()
    pairs = drv.shard(list(zip(img_paths, mask_paths)))

    def produce():
        for ip, mp in pairs:
            img = drv.read_rgb(ip)
            x = torch.from_numpy(preprocessing_fn(img).transpose(2, 0, 1)).float()[None].pin_memory()
            prob = drv.predict_probs(model, transforms, x.to(dev, non_blocking=True))
            mask = drv.read_mask(mp, 50)
            yield drv.scored(prob[0].contiguous(), mask), mask, str(ip).split("/")[-1]

    _finish(drv.CachedPredictions(produce), logdir, config, as_float=False)


def tta_patches(logdir, config, args):
    test_img_dir = config["test_img_path"]
    test_mask_dir = config["test_mask_path"] / lesion_dict[config["lesion_type"]].dir_name
    ALL_MASKS = sorted(test_mask_dir.glob("*.*"))
    model = drv.build_model(config, logdir, args)
    _, mean, std = archs.get_preprocessing_fn(dataset_name=config["dataset_name"], grayscale=config["gray"])
    if config["gray"]:
        raise NotImplementedError("gray=True inputs are outside the B200 hot path (3-channel stem)")
    transforms = drv.tta_transforms(args)
    resize_size = config["scale_size"]
    blend = drv.tile_blend_mode(config)

    def load(mask_path):
        return drv.read_rgb(test_img_dir / mask_path.name), drv.read_mask(mask_path, 50)

    rank, world_size = partition.world()
    if world_size > 1 and blend == "overwrite":          # (image, tile) units over the ranks, as in tta.tta_patches
        produce = drv.partitioned_producer(model, transforms, ALL_MASKS, load, resize_size, mean, std)
    else:
        MY_MASKS = drv.shard(ALL_MASKS)

        def produce():
            loaded = drv.prefetched([(lambda m=m: load(m)) for m in MY_MASKS])
            for mask_path, (image, gt_mask) in zip(MY_MASKS, loaded):
                pred, _ = drv.infer_image_host(model, transforms, torch.from_numpy(image), torch.from_numpy(gt_mask),
                                               resize_size, mean, std, blend=blend)
                yield pred, gt_mask, mask_path.name

    _finish(drv.CachedPredictions(produce), logdir, config, as_float=True)
