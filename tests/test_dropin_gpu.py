"""Drop-in drivers end to end, from files on disk to files on disk: ``tta_patches(logdir, config, args)``
(src/main/tta.py:150-238) and its vessel twin (tta_vessel.py:138-229) with the reference's config keys,
checkpoint location and output layout, followed by ``export_result`` (stat_result.py) on what they wrote --
the exact sequence of pipeline.py:63-107.  The checker is the oracle pipeline (oracle/pipeline.py +
oracle/scoring.py) run on the same files."""
import os
from pathlib import Path

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

from eyediseasesegmentation_b200 import stat_result, tta as eds_tta, tta_vessel as eds_tta_vessel  # noqa: E402
from oracle import nets, pipeline, scoring  # noqa: E402
import helpers  # noqa: E402

S = 128
NAME, CFG = "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)


def _fundus(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[:h, :w]
    img[(yy - h / 2) ** 2 + (xx - w / 2) ** 2 > (0.48 * min(h, w)) ** 2] = 0
    return img


def _checkpoint(tmp_path, sd):
    logdir = tmp_path / "models" / "IDRiD" / "EX" / "exp1"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "best.pth")
    return logdir


@pytest.fixture()
def fp32_mode(monkeypatch):
    monkeypatch.setenv("EDS_PRECISION", "fp32")


@pytest.mark.parametrize("gray", [False, True])
def test_tta_patches_from_disk_to_disk(tmp_path, fp32_mode, gray):
    """gray=True on this path (tta.py:166) only replaces the per-channel statistics by their luma-weighted scalars;
    the window is still read as RGB."""
    model = helpers.build_product_model(NAME, CFG)
    sd = model.state_dict()
    logdir = _checkpoint(tmp_path, sd)
    img_dir = tmp_path / "data" / "images"
    mask_root = tmp_path / "data" / "masks"
    mask_dir = mask_root / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(11)
    shapes = [(300, 420), (280, 300), (330, 290)]
    for i, (h, w) in enumerate(shapes):
        Image.fromarray(_fundus(h, w, 20 + i)).save(img_dir / f"IDRiD_{i:02d}.jpg", quality=95)
        gt = (np.kron(rng.random((h // 10 + 1, w // 10 + 1)) < 0.08, np.ones((10, 10)))[:h, :w] * 255).astype(np.uint8)
        if i == 2:
            gt[:] = 0                                     # an image without positives (aucpr.py:22)
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{i:02d}_EX.tif")
    out_dir = tmp_path / "outputs"
    config = {"dataset_name": "IDRiD", "lesion_type": "EX", "gray": gray, "scale_size": S, "val_batch_size": 2,
              "model_name": NAME, "model_params": dict(CFG), "test_img_path": img_dir, "test_mask_path": mask_root,
              "out_dir": str(out_dir), "data_type": "tile"}
    eds_tta.tta_patches(str(logdir), config, {"best": "true", "tta": "d4", "createprob": "false", "optim_thres": 0})

    # ---- oracle on the same files
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    if gray:                                               # archs/__init__.py:84-86
        mean = mean[0] * 0.2989 + mean[1] * 0.5870 + mean[2] * 0.1140
        std = std[0] * 0.2989 + std[1] * 0.5870 + std[2] * 0.1140
    items = []
    for mp in sorted(mask_dir.glob("*.*")):
        image = np.asarray(Image.open(img_dir / mp.name.replace("_EX.tif", ".jpg")).convert("RGB")).astype("uint8")
        gt = (np.asarray(Image.open(mp).convert("L")) > 0).astype(np.uint8)
        pred = pipeline.tiled_probability_map(image, lambda t: nets.unetplusplus_forward(sd, t), S, mean, std, "d4")
        items.append((pred, gt, mp.name))
    t3 = scoring.pr_curve(items)["thresholds"][2]          # the one tta.py:221,226 uses
    written = out_dir / "IDRiD" / "tta" / "EX" / "exp1"
    assert sorted(p.name for p in written.iterdir()) == [f"IDRiD_{i:02d}.jpg" for i in range(len(shapes))]
    for pred, _, name in items:
        got = np.asarray(Image.open(written / name.replace("_EX.tif", ".jpg")).convert("L")) > 127
        want = pred > t3
        if want.all():                                     # save_output min-max rescales: a constant mask is saved black
            want = np.zeros_like(want)
        assert got.shape == want.shape
        # JPEG ringing and probabilities within 1e-4 of the threshold may flip isolated pixels
        assert np.mean(got != want) < 5e-3, name

    # ---- pipeline.py:107: the per-image metrics of the masks just written
    stat_result.export_result("EX/exp1", config)
    rows = helpers.read_stat_csvs(out_dir / "IDRiD" / "result_assessment" / "EX" / "exp1")
    assert set(rows["dice"]) == {f"IDRiD_{i:02d}_EX.tif" for i in range(len(shapes))} | {"Avg:"}


def test_vessel_test_tta_from_disk_to_disk(tmp_path, fp32_mode):
    """tta_vessel.test_tta (tta_vessel.py:55-136): whole pre-padded square images through the PROPOSED
    network (base_dim 8 -> 256^2), D4 TTA, ROC scoring, masks written under the image's own name;
    then stat_result_vessel.export_result on them."""
    from eyediseasesegmentation_b200 import stat_result_vessel
    name, cfg = "unetplusplusstar", helpers.star_cfg(8)
    model = helpers.build_product_model(name, cfg)
    sd = model.state_dict()
    logdir = tmp_path / "models" / "DRIVE" / "Vessel_DRIVE" / "vexp"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "last.pth")
    img_dir, mask_dir = tmp_path / "vdata" / "images", tmp_path / "vdata" / "masks"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(12)
    for i in range(3):
        Image.fromarray(_fundus(256, 256, 40 + i)).save(img_dir / f"{i:02d}_test.jpg", quality=95)
        gt = (np.kron(rng.random((32, 32)) < 0.15, np.ones((8, 8))) * 255).astype(np.uint8)
        Image.fromarray(gt, "L").save(mask_dir / f"{i:02d}_test.jpg", quality=100)
    out_dir = tmp_path / "voutputs"
    config = {"dataset_name": "DRIVE", "lesion_type": "Vessel_DRIVE", "gray": False, "scale_size": 256,
              "val_batch_size": 1, "model_name": name, "model_params": dict(cfg), "test_img_path": img_dir,
              "test_mask_path": mask_dir, "out_dir": str(out_dir), "data_type": "all"}
    eds_tta_vessel.test_tta(str(logdir), config, {"best": "false", "tta": "d4"})

    mean, std = pipeline.DATASET_STATS["IDRiD"]          # tta_vessel.py:73 passes dataset_name=None
    items = []
    for ip in sorted(img_dir.glob("*.jpg")):
        image = np.asarray(Image.open(ip).convert("RGB")).astype("uint8")
        gt = (np.asarray(Image.open(mask_dir / ip.name).convert("L")) > 50).astype(np.uint8)
        x = torch.from_numpy(pipeline.preprocess(image, mean, std).transpose(2, 0, 1)).float()[None]
        with torch.no_grad():
            logit = nets.tta_mean_logits(lambda t: nets.unetplusplusstar_forward(sd, t, 8), x, "d4")[0, 0]
        items.append((torch.sigmoid(logit).numpy(), gt, ip.name))
    t = scoring.roc_curve(items)["threshold"]
    written = out_dir / "DRIVE" / "tta" / "Vessel_DRIVE" / "vexp"
    assert sorted(p.name for p in written.iterdir()) == sorted(n for _, _, n in items)
    for pred, _, n in items:
        got = np.asarray(Image.open(written / n).convert("L")) > 127
        assert np.mean(got != (pred > t)) < 5e-3, n
    config["test_mask_path"] = mask_dir
    stat_result_vessel.export_result("Vessel_DRIVE/vexp", config)
    rows = helpers.read_stat_csvs(out_dir / "DRIVE" / "result_assessment" / "Vessel_DRIVE" / "vexp")
    assert set(rows["sn"]) == {n for _, _, n in items} | {"Avg:"}


def test_lesion_test_tta_whole_image_from_disk_to_disk(tmp_path, fp32_mode):
    """tta.test_tta (tta.py:56-148), BASELINE config 1 network (smp.Unet / resnet34): LongestMaxSize(1024) +
    centred pad on the host, batch of 2 through hflip TTA, sigmoid, centre crop + bilinear resize back to the
    original size, PR scoring at the original size, masks written under the image name."""
    import cv2
    name, cfg = "Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1)
    model = helpers.build_product_model(name, cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["segmentation_head.0.weight"] *= 6.0            # spread the random-init logits across the thresholds
    sd["segmentation_head.0.bias"] -= 1.0
    logdir = tmp_path / "models" / "IDRiD" / "EX" / "whole"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "best.pth")
    img_dir = tmp_path / "wdata" / "images"
    mask_root = tmp_path / "wdata" / "masks"
    mask_dir = mask_root / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    H0, W0 = 356, 536                                  # 1/8 of IDRiD's 2848 x 4288: same aspect, same pad geometry
    S = 1024
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    scale = S / max(H0, W0)
    nh, nw = int(round(H0 * scale)), int(round(W0 * scale))
    top, left = int((S - nh) / 2.0), int((S - nw) / 2.0)

    # oracle first: its probability maps also define ground truths the random-init network "knows"
    preds = []
    for i in range(2):
        Image.fromarray(_fundus(H0, W0, 60 + i)).save(img_dir / f"IDRiD_{i:02d}.jpg", quality=95)
        image = np.asarray(Image.open(img_dir / f"IDRiD_{i:02d}.jpg").convert("RGB")).astype("uint8")
        small = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_LINEAR)
        padded = np.zeros((S, S, 3), dtype=np.uint8)
        padded[top:top + nh, left:left + nw] = small
        x = torch.from_numpy(pipeline.preprocess(padded, mean, std).transpose(2, 0, 1)).float()[None]
        with torch.no_grad():
            logit = nets.tta_mean_logits(lambda t: nets.unet_forward(sd, t), x, "hflip")[0, 0]
        pred = pipeline.whole_image_probability(torch.sigmoid(logit).numpy(), (nh, nw), (H0, W0))
        preds.append(pred)
        rng = np.random.default_rng(70 + i)
        top30 = np.zeros(pred.size, dtype=bool)
        top30[np.argsort(pred.ravel(), kind="stable")[-int(0.3 * pred.size):]] = True      # rank based: ties cannot empty it
        gt = (top30.reshape(pred.shape) ^ (rng.random(pred.shape) < 0.05)).astype(np.uint8) * 255
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{i:02d}_EX.tif")

    out_dir = tmp_path / "woutputs"
    config = {"dataset_name": "IDRiD", "lesion_type": "EX", "gray": False, "scale_size": S, "val_batch_size": 2,
              "model_name": name, "model_params": dict(cfg), "test_img_path": img_dir, "test_mask_path": mask_root,
              "out_dir": str(out_dir), "data_type": "all"}
    eds_tta.test_tta(str(logdir), config, {"best": "true", "tta": "hflip"})

    items = []
    for i in range(2):
        m = (np.asarray(Image.open(mask_dir / f"IDRiD_{i:02d}_EX.tif").convert("L")) > 50).astype(np.uint8)
        ms = cv2.resize(m, (nw, nh), interpolation=cv2.INTER_NEAREST)
        mp = np.zeros((S, S), dtype=np.uint8)
        mp[top:top + nh, left:left + nw] = ms
        gt = cv2.resize(mp[(S - nh) // 2:(S - nh) // 2 + nh, (S - nw) // 2:(S - nw) // 2 + nw], (W0, H0),
                        interpolation=cv2.INTER_LINEAR)
        items.append((preds[i], gt, f"IDRiD_{i:02d}.jpg"))
    t3 = scoring.pr_curve(items)["thresholds"][2]
    written = out_dir / "IDRiD" / "tta" / "EX" / "whole"
    assert sorted(p.name for p in written.iterdir()) == [n for _, _, n in items]
    for pred, _, n in items:
        got = np.asarray(Image.open(written / n).convert("L")) > 127
        assert got.shape == pred.shape
        want = pred > t3
        if want.all():                                 # save_output min-max rescales: a constant mask is saved black
            want = np.zeros_like(want)
        assert 0.02 < want.mean() < 0.98, "degenerate test image"
        assert np.mean(got != want) < 5e-3, n


def test_ensemble_predict_from_disk_to_disk(tmp_path, fp32_mode):
    """ensemble.predict (reference ensemble.py:64-125): two different architectures, each read from its own
    ``config.json`` + ``checkpoints/best.pth``, D4 TTA each, per-image mean of the sigmoids at S x S, PR scoring,
    masks written under the image name.  Checker: oracle.pipeline.ensemble_probability on the same files."""
    import json
    import cv2
    from eyediseasesegmentation_b200 import ensemble as eds_ensemble
    S = 128
    specs = [("unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1),
              lambda sd: (lambda t: nets.unetplusplus_forward(sd, t))),
             ("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1),
              lambda sd: (lambda t: nets.unet_forward(sd, t)))]
    logdirs, oracle_nets = [], []
    for i, (name, cfg, mk) in enumerate(specs):
        model = helpers.build_product_model(name, cfg, seed=1999 + i)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        sd["segmentation_head.0.weight"] *= 4.0
        logdir = tmp_path / "models" / "IDRiD" / "EX" / f"run{i}"
        (logdir / "checkpoints").mkdir(parents=True)
        torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "best.pth")
        with open(logdir / "config.json", "w") as j:
            json.dump({"model_name": name, "model_params": cfg}, j)
        logdirs.append(logdir)
        oracle_nets.append(mk(sd))

    img_dir = tmp_path / "edata" / "images"
    mask_root = tmp_path / "edata" / "masks"
    mask_dir = mask_root / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    H0, W0 = 178, 268
    scale = S / max(H0, W0)
    nh, nw = int(round(H0 * scale)), int(round(W0 * scale))
    top, left = int((S - nh) / 2.0), int((S - nw) / 2.0)
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    rng = np.random.default_rng(5)
    items = []
    for i in range(3):
        Image.fromarray(_fundus(H0, W0, 80 + i)).save(img_dir / f"IDRiD_{i:02d}.jpg", quality=95)
        gt = (np.kron(rng.random((H0 // 8 + 1, W0 // 8 + 1)) < 0.2, np.ones((8, 8)))[:H0, :W0] * 255).astype(np.uint8)
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{i:02d}_EX.tif")
        image = np.asarray(Image.open(img_dir / f"IDRiD_{i:02d}.jpg").convert("RGB")).astype("uint8")
        padded = np.zeros((S, S, 3), dtype=np.uint8)
        padded[top:top + nh, left:left + nw] = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_LINEAR)
        x = torch.from_numpy(pipeline.preprocess(padded, mean, std).transpose(2, 0, 1)).float()[None]
        pred = pipeline.ensemble_probability(oracle_nets, x)
        m = (np.asarray(Image.open(mask_dir / f"IDRiD_{i:02d}_EX.tif").convert("L")) > 50).astype(np.uint8)
        mp = np.zeros((S, S), dtype=np.uint8)
        mp[top:top + nh, left:left + nw] = cv2.resize(m, (nw, nh), interpolation=cv2.INTER_NEAREST)
        items.append((pred, mp, f"IDRiD_{i:02d}.jpg"))

    out_dir = tmp_path / "eoutputs"
    config = {"dataset_name": "IDRiD", "lesion_type": "EX", "scale_size": S, "out_dir": str(out_dir),
              "test_img_path": img_dir, "test_mask_path": mask_root}
    got_auc = eds_ensemble.predict(config, logdirs, "ensemble_1")

    want_auc = scoring.get_auc(items)
    assert abs(got_auc - want_auc) < 1e-3                      # BASELINE: AUC-PR within 1e-3
    t1 = scoring.pr_curve(items)["thresholds"][0]
    written = out_dir / "IDRiD" / "tta" / "EX" / "ensemble_1"
    assert sorted(p.name for p in written.iterdir()) == [n for _, _, n in items]
    for pred, _, n in items:
        got = np.asarray(Image.open(written / n).convert("L")) > 127
        want = pred > t1
        if want.all():
            want = np.zeros_like(want)
        assert got.shape == (S, S)
        assert np.mean(got != want) < 5e-3, n
