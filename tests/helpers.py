"""Shared test helpers (synthetic weights, error metrics, oracle dispatch)."""
import torch

from oracle import nets


def star_cfg(base_dim):
    """config.py:82-93 of the reference with the scratch encoder (no checkpoint download)."""
    return dict(classes=1, decoder_attention_type="scse", decoder_use_batchnorm=True, base_dim=base_dim,
                encoder_depth=5, encoder_name="BoTSER50_Axial_scratch", deep_supervision=False,
                drop_block_prob=0.0, clf_head=False)


def randomize_bn(model, seed):
    """Non-trivial BatchNorm statistics so that BN folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in model.state_dict().items():
            if name.endswith("running_mean"):
                t.copy_(torch.randn(t.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                t.copy_(torch.rand(t.shape, generator=g) * 0.5 + 0.75)
        for name, p in model.named_parameters():
            base = name.rsplit(".", 1)[0]
            if (base + ".running_mean") in model.state_dict():
                if name.endswith(".weight"):
                    p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.75)
                elif name.endswith(".bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return model


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def oracle_forward(name, cfg, state_dict, x, features=False, device="cpu"):
    sd = {k: v.detach().to(device) for k, v in state_dict.items()}
    x = x.to(device)
    with torch.no_grad():
        if name == "unetplusplusstar":
            out = nets.unetplusplusstar_forward(sd, x, int(cfg["base_dim"]), return_features=features)
        elif name == "unetplusplus_deepsup":
            out = nets.unetplusplus_forward(sd, x, return_features=features)
        elif name == "Unet":
            out = nets.unet_forward(sd, x, return_features=features)
        else:
            raise KeyError(name)
    if features:
        return out[0].cpu(), [f.cpu() for f in out[1]]
    return out.cpu()


# ------------------------------------------------------------------ synthetic data
NET_GOLDEN_CASES = [
    # key, registry name, params, input size, batch
    ("star_bd8", "unetplusplusstar", star_cfg(8), 256, 1),
    ("upp_se50_scse", "unetplusplus_deepsup", dict(encoder_name="se_resnet50", encoder_weights=None, classes=1,
                                                   decoder_attention_type="scse", deep_supervision=True), 128, 1),
    ("upp_r34", "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 128, 2),
    ("unet_r34", "Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 128, 2),
]


def build_product_model(name, cfg, seed=1999):
    """Product-side model with deterministic synthetic weights (same on every machine)."""
    from eyediseasesegmentation_b200 import archs
    import copy
    cfg = copy.deepcopy(cfg)
    torch.manual_seed(seed)
    model = archs.Unet(**cfg) if name == "Unet" else archs.get_model(name, cfg, training=False)
    randomize_bn(model, seed + 1)
    return model.eval()


def golden_input(batch, size, seed=7):
    return torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(seed))


def synth_scoring_case(seed, shape=(96, 160), n_images=5):
    """Probability maps + {0,1} masks with a trained-model-like score distribution, exact
    threshold hits, flat background and one image without positives (aucpr.py:22)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    th32 = np.array([0, 0.00001, 0.0001, 0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.99, 0.999,
                     0.9999, 0.99999, 1], dtype=np.float32)
    items = []
    for i in range(n_images):
        logit = rng.normal(-4.0, 3.0, size=shape).astype(np.float32)
        gt = (rng.random(shape) < 1 / (1 + np.exp(-(logit * 0.8 + rng.normal(-1, 1.5, shape))))).astype(np.uint8)
        prob = (1 / (1 + np.exp(-logit))).astype(np.float32)
        prob[:8, :] = prob[8, 0]                                   # flat background rows
        flat = prob.reshape(-1)
        flat[1000:1019] = th32                                      # scores exactly on the thresholds
        flat[1019:1038] = np.nextafter(th32, np.float32(2))         # ... one ulp above
        flat[1038:1057] = np.nextafter(th32, np.float32(-1))        # ... one ulp below
        if i == n_images - 1:
            gt[:] = 0
        items.append((prob, gt, f"img{i}"))
    return items


def make_stat_case(root, seed=0, n_images=6, shape=(120, 168)):
    """Synthetic mask set for export_result (stat_result.py): ground truth under
    <root>/masks/<lesion dir>/IDRiD_NN_EX.tif, saved predictions under <root>/out/IDRiD/tta/EX/exp/IDRiD_NN.jpg
    (lesion layout) and <root>/vmasks/NN_manual1.png with predictions of the same name (vessel layout).
    Grey levels around the `> 50` threshold, an image without positives, one without predictions,
    one where both are empty and one all-positive ground truth exercise every branch of :55-79."""
    import numpy as np
    from pathlib import Path
    from PIL import Image
    from eyediseasesegmentation_b200.util import lesion_dict
    rng = np.random.default_rng(seed)
    root = Path(root)
    gt_dir = root / "masks" / lesion_dict["EX"].dir_name
    pred_dir = root / "out" / "IDRiD" / "tta" / "EX" / "exp"
    vgt_dir = root / "vmasks"
    vpred_dir = root / "out" / "DRIVE" / "tta" / "vexp"
    for d in (gt_dir, pred_dir, vgt_dir, vpred_dir):
        d.mkdir(parents=True, exist_ok=True)
    levels = np.array([0, 30, 50, 51, 128, 255], dtype=np.uint8)
    for i in range(n_images):
        gt = levels[rng.integers(0, len(levels), size=shape)]
        blob = rng.random((shape[0] // 8, shape[1] // 8)) < 0.3
        pred = (np.kron(blob, np.ones((8, 8), dtype=np.uint8)) * 255).astype(np.uint8)
        if i == 1:
            gt[:] = 0
        if i == 2:
            pred[:] = 0
        if i == 3:
            gt[:] = 0
            pred[:] = 0
        if i == 4:
            gt[:] = 255
        Image.fromarray(gt, "L").save(gt_dir / f"IDRiD_{i:02d}_EX.tif")
        Image.fromarray(pred, "L").save(pred_dir / f"IDRiD_{i:02d}.jpg", quality=95)
        Image.fromarray(gt, "L").save(vgt_dir / f"{i:02d}_manual1.png")
        Image.fromarray(pred, "L").save(vpred_dir / f"{i:02d}_manual1.png")
    lesion_cfg = {"test_mask_path": root / "masks", "lesion_type": "EX", "out_dir": str(root / "out"),
                  "dataset_name": "IDRiD"}
    vessel_cfg = {"test_mask_path": vgt_dir, "lesion_type": "Vessel_DRIVE", "out_dir": str(root / "out"),
                  "dataset_name": "DRIVE"}
    return lesion_cfg, vessel_cfg


def read_stat_csvs(directory):
    """{metric: {row name: text of the value}} of the five CSVs export_result writes."""
    import os
    out = {}
    for name in ("sn", "ppv", "sp", "iou", "dice"):
        rows = {}
        for line in open(os.path.join(directory, name + ".csv")).read().splitlines():
            key, val = line.rsplit(",", 1)
            rows[key] = val
        out[name] = rows
    return out


def assert_stat_csvs_equal(got, want):
    """Per-image rows byte for byte; the 'Avg:' row to 1e-12 (np.mean follows os.listdir order, which
    is a property of the file system, not of the code)."""
    for name in want:
        assert set(got[name]) == set(want[name]), name
        for key, val in want[name].items():
            if key == "Avg:":
                assert abs(float(got[name][key]) - float(val)) <= 1e-12 * max(1.0, abs(float(val))), (name, key)
            else:
                assert got[name][key] == val, (name, key, got[name][key], val)


def make_pad_case(root, seed=0):
    """Tiny unpadded vessel-style files for pad_img.pad: RGB images as PNG and single-channel label PNGs with grey
    levels on both sides of the 127 threshold; odd and even deltas (DRIVE 584x565 -> 608 has an odd width delta)."""
    import numpy as np
    from pathlib import Path
    from PIL import Image
    rng = np.random.default_rng(seed)
    root = Path(root)
    (root / "images").mkdir(parents=True, exist_ok=True)
    (root / "labels").mkdir(parents=True, exist_ok=True)
    shapes = {"a.png": (30, 27), "b.png": (29, 32), "c.png": (32, 32)}
    for name, (h, w) in shapes.items():
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(root / "images" / name)
        levels = np.array([0, 1, 126, 127, 128, 200, 255], dtype=np.uint8)
        Image.fromarray(levels[rng.integers(0, len(levels), size=(h, w))], "L").save(root / "labels" / name)
    return root / "images", root / "labels", 32


def read_pad_outputs(directory):
    """{file name: array as cv2 reads it back (BGR for images, 2-D for labels)} of a pad() output folder."""
    import cv2
    import os
    return {f: cv2.imread(os.path.join(directory, f), cv2.IMREAD_UNCHANGED) for f in sorted(os.listdir(directory))}


# ------------------------------------------------------------------ ensemble (reference ensemble.py)
ENSEMBLE_S = 64
ENSEMBLE_SPECS = [("unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)),
                  ("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1))]
ENSEMBLE_IMG_DIR = "data/raw/IDRiD/1. Original Images/b. Testing Set"                   # ensemble.py:65
ENSEMBLE_MASK_DIR = "data/raw/IDRiD/2. All Segmentation Groundtruths/b. Testing Set"    # ensemble.py:66


def ensemble_state_dicts():
    """The two member models of the ensemble fixture: deterministic product-initialised weights, segmentation head
    scaled up so that the probabilities spread over (0, 1)."""
    out = []
    for i, (name, cfg) in enumerate(ENSEMBLE_SPECS):
        sd = {k: v.clone() for k, v in build_product_model(name, cfg, seed=1999 + i).state_dict().items()}
        sd["segmentation_head.0.weight"] *= 8.0
        out.append(sd)
    return out


def ensemble_oracle_nets(state_dicts):
    return [lambda t, sd=state_dicts[0]: nets.unetplusplus_forward(sd, t),
            lambda t, sd=state_dicts[1]: nets.unet_forward(sd, t)]


def make_ensemble_case(root, gts=None, n_images=3, seed=11):
    """Under ``root``: the IDRiD test folders at the paths ensemble.py hard-codes (S x S JPEG images + 'EX' label
    TIFFs) and one log directory per member model (config.json + checkpoints/best.pth).  ``gts`` [n, S, S] {0,1}:
    the label maps to write (the committed fixture's, which were derived from the ensemble's own probabilities so
    that the PR curve and the masks are not degenerate); random blocks without it.  -> (config, logdirs)"""
    import json
    import numpy as np
    from pathlib import Path
    from PIL import Image
    root = Path(root)
    S = ENSEMBLE_S
    img_dir = root / ENSEMBLE_IMG_DIR
    mask_dir = root / ENSEMBLE_MASK_DIR / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(seed)
    for i in range(n_images):
        Image.fromarray(rng.integers(0, 256, size=(S, S, 3), dtype=np.uint8)).save(img_dir / f"IDRiD_{55 + i}.jpg", quality=95)
        gt = (np.kron(rng.random((S // 8, S // 8)) < 0.25, np.ones((8, 8))) * 255).astype(np.uint8)
        if gts is not None:
            gt = (np.asarray(gts[i]) > 0).astype(np.uint8) * 255
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{55 + i}_EX.tif")
    logdirs = []
    for i, ((name, cfg), sd) in enumerate(zip(ENSEMBLE_SPECS, ensemble_state_dicts())):
        logdir = root / "models" / "IDRiD" / "EX" / f"run{i}"
        (logdir / "checkpoints").mkdir(parents=True)
        torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "best.pth")
        with open(logdir / "config.json", "w") as j:
            json.dump({"model_name": name, "model_params": cfg}, j)
        logdirs.append(logdir)
    config = {"lesion_type": "EX", "dataset_name": "IDRiD", "scale_size": S, "out_dir": str(root / "outputs")}
    return config, logdirs


def run_reference_ensemble(root, gts=None):
    """Runs the reference's own ensemble.predict (oracle/ref_loader.load_ensemble) on make_ensemble_case(root) with
    the working directory at ``root`` (its dataset paths are relative).  -> dict of arrays, file-name order."""
    import os
    import numpy as np
    from pathlib import Path
    from PIL import Image
    from oracle import ref_loader
    config, logdirs = make_ensemble_case(root, gts=gts)
    ens, cap = ref_loader.load_ensemble()
    # in-process loading: ensemble.py:80 asks for two worker processes, and fork() from a multi-threaded test process
    # can deadlock; the loader only feeds images (still shuffled), the arithmetic under test is untouched
    real_loader = torch.utils.data.DataLoader
    ens.DataLoader = lambda ds, **kw: real_loader(ds, **{**kw, "num_workers": 0, "pin_memory": False})
    cwd = os.getcwd()
    os.chdir(root)
    try:
        # the file names come back from the reference's save loop (shuffled loader order, ensemble.py:80)
        ens.predict(config, [Path(os.path.relpath(d, root)) for d in logdirs], "ensemble_1")
    finally:
        os.chdir(cwd)
    names = list(cap["masks"])                       # save order == prediction order (ensemble.py:112)
    order = sorted(range(len(names)), key=lambda i: names[i])
    img_dir = Path(root) / ENSEMBLE_IMG_DIR
    return {"names": np.array([names[i] for i in order]),
            "images": np.stack([np.asarray(Image.open(img_dir / names[i]).convert("RGB")) for i in order]),
            "gts": np.stack([np.asarray(cap["gts"][i]).reshape(ENSEMBLE_S, ENSEMBLE_S) for i in order]).astype(np.uint8),
            "preds": np.stack([cap["preds"][i] for i in order]).astype(np.float32),
            "masks": np.stack([cap["masks"][names[i]] for i in order]).astype(np.uint8),
            "auc": np.float64(cap["auc"]), "thresholds": np.array(cap["thresholds"], dtype=np.float64)}


# ------------------------------------------------------------------ tta.py drivers (reference tta_patches / test_tta)
TTA_CASES = {
    # name: registry name, params, scale_size, TTA alias, original image shapes, data_type
    "patches": ("unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 64, "d4",
                [(150, 200), (140, 131), (128, 170)]),
    "whole": ("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 1024, "hflip",
              [(89, 134), (89, 134)]),
    # tta_vessel.test_tta: whole pre-padded squares through the proposed network (base_dim 8 <-> 256^2), ROC scoring
    "vessel": ("unetplusplusstar", star_cfg(8), 256, "d4", [(256, 256), (256, 256)]),
    # tta_vessel.tta_patches: sliding window over unpadded vessel images, DRIVE statistics, ROC scoring
    "vpatches": ("unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 64, "d4",
                 [(150, 180), (131, 140)]),
}
VESSEL_CASES = ("vessel", "vpatches")


def tta_case_state_dict(case):
    name, cfg = TTA_CASES[case][:2]
    sd = {k: v.clone() for k, v in build_product_model(name, cfg, seed=2001).state_dict().items()}
    # fixed head calibration (measured once on the case's first image): logits ~ mean 0, std 2, so that the
    # probabilities of the random-init network spread over the threshold list
    scale, bias = {"patches": (5.0, -5.9), "whole": (1.0, 3.6), "vessel": VESSEL_HEAD, "vpatches": VPATCHES_HEAD}[case]
    sd["segmentation_head.0.weight"] *= scale
    sd["segmentation_head.0.bias"] += bias
    return sd


VESSEL_HEAD = (12.0, 8.0)
VPATCHES_HEAD = (6.0, -4.75)


def tta_case_oracle_net(case, sd):
    name, cfg = TTA_CASES[case][:2]
    if name == "Unet":
        return lambda t: nets.unet_forward(sd, t)
    if name == "unetplusplusstar":
        return lambda t: nets.unetplusplusstar_forward(sd, t, int(cfg["base_dim"]))
    return lambda t: nets.unetplusplus_forward(sd, t)


def _make_vessel_case(root, case, jpegs, gts, mask_jpegs, seed):
    """tta_vessel layout: images and labels are both ``NN_test.jpg`` (get_datapath globs *.jpg for 'Vessel_*'), the
    labels are read with ``> 50`` (lesion_dataset.py:127), the checkpoint is ``last.pth``."""
    import io
    import numpy as np
    from pathlib import Path
    from PIL import Image
    name, cfg, S, alias, shapes = TTA_CASES[case]
    root = Path(root)
    img_dir, mask_dir = root / "vimages", root / "vmasks"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(seed)

    def jpeg(a, q):
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="JPEG", quality=q)
        return buf.getvalue()

    for i, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        yy, xx = np.mgrid[:h, :w]
        img[(yy - h / 2) ** 2 + (xx - w / 2) ** 2 > (0.48 * min(h, w)) ** 2] = 0
        gt = (np.kron(rng.random((h // 8 + 1, w // 8 + 1)) < 0.15, np.ones((8, 8)))[:h, :w] * 255).astype(np.uint8)
        if gts is not None:
            gt = (np.asarray(gts[i]) > 0).astype(np.uint8) * 255
        (img_dir / f"{i:02d}_test.jpg").write_bytes(bytes(np.asarray(jpegs[i], dtype=np.uint8)) if jpegs is not None
                                                    else jpeg(img, 95))
        (mask_dir / f"{i:02d}_test.jpg").write_bytes(bytes(np.asarray(mask_jpegs[i], dtype=np.uint8))
                                                     if mask_jpegs is not None else jpeg(gt, 100))
    logdir = root / "models" / "DRIVE" / "Vessel_DRIVE" / f"{case}_exp"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": tta_case_state_dict(case)}, logdir / "checkpoints" / "last.pth")
    config = {"dataset_name": "DRIVE", "lesion_type": "Vessel_DRIVE", "gray": False, "scale_size": S,
              "val_batch_size": 1, "model_name": name, "model_params": dict(cfg), "test_img_path": img_dir,
              "test_mask_path": mask_dir, "out_dir": str(root / "outputs"),
              "data_type": "all" if case == "vessel" else "tile"}
    return logdir, config, {"best": "false", "tta": alias}


def make_tta_case(root, case, jpegs=None, gts=None, seed=31, mask_jpegs=None):
    """Under ``root``: ``images/IDRiD_NN.jpg``, ``masks/3. Hard Exudates/IDRiD_NN_EX.tif`` and the log directory of one
    model.  ``jpegs`` (list of uint8 byte arrays) / ``gts`` ([H,W] {0,1} arrays): the committed fixture's files, written
    back verbatim (a re-encoded JPEG would not be the same image); seeded noise discs / random blocks without them.
    -> (logdir, config, args)"""
    import io
    import json
    import numpy as np
    from pathlib import Path
    from PIL import Image
    if case in VESSEL_CASES:
        return _make_vessel_case(root, case, jpegs, gts, mask_jpegs, seed)
    name, cfg, S, alias, shapes = TTA_CASES[case]
    root = Path(root)
    img_dir, mask_root = root / "images", root / "masks"
    mask_dir = mask_root / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(seed)
    for i, (h, w) in enumerate(shapes):
        if jpegs is not None:
            (img_dir / f"IDRiD_{i:02d}.jpg").write_bytes(bytes(np.asarray(jpegs[i], dtype=np.uint8)))
        else:
            img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            yy, xx = np.mgrid[:h, :w]
            img[(yy - h / 2) ** 2 + (xx - w / 2) ** 2 > (0.55 * min(h, w)) ** 2] = 0
            buf = io.BytesIO()
            Image.fromarray(img).save(buf, format="JPEG", quality=95)
            (img_dir / f"IDRiD_{i:02d}.jpg").write_bytes(buf.getvalue())
        gt = (np.kron(rng.random((h // 8 + 1, w // 8 + 1)) < 0.2, np.ones((8, 8)))[:h, :w] * 255).astype(np.uint8)
        if gts is not None:
            gt = (np.asarray(gts[i]) > 0).astype(np.uint8) * 255
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{i:02d}_EX.tif")
    logdir = root / "models" / "IDRiD" / "EX" / f"{case}_exp"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": tta_case_state_dict(case)}, logdir / "checkpoints" / "best.pth")
    config = {"dataset_name": "IDRiD", "lesion_type": "EX", "gray": False, "scale_size": S, "val_batch_size": 2,
              "model_name": name, "model_params": dict(cfg), "test_img_path": img_dir, "test_mask_path": mask_root,
              "out_dir": str(root / "outputs"), "data_type": "tile" if case == "patches" else "all"}
    args = {"best": "true", "tta": alias, "createprob": "false", "optim_thres": 0}
    return logdir, config, args


def run_reference_tta(root, case, jpegs=None, gts=None, mask_jpegs=None):
    """Runs the reference's own tta.tta_patches / tta.test_tta (oracle/ref_loader.load_tta) on make_tta_case.
    -> dict of arrays (per image, in file-name order: JPEG bytes, label map as stored, probability map and label map
    as scored, the mask handed to save_output) + auc + thresholds."""
    import copy
    import numpy as np
    from pathlib import Path
    from oracle import ref_loader
    logdir, config, args = make_tta_case(root, case, jpegs=jpegs, gts=gts, mask_jpegs=mask_jpegs)
    mod, cap = ref_loader.load_tta("tta_vessel" if case in VESSEL_CASES else "tta")
    (mod.tta_patches if case in ("patches", "vpatches") else mod.test_tta)(str(logdir), copy.deepcopy(config), dict(args))
    key = (lambda n: n.replace("_EX.tif", ".jpg"))
    items = sorted(cap["items"], key=lambda it: key(it[2]))
    out = {"auc": np.float64(cap["auc"]), "thresholds": np.array(cap["thresholds"], dtype=np.float64),
           "names": np.array([key(n) for _, _, n in items])}
    for i, (pred, gt, n) in enumerate(items):
        out[f"jpeg{i}"] = np.frombuffer((Path(config["test_img_path"]) / key(n)).read_bytes(), dtype=np.uint8)
        if case in VESSEL_CASES:
            out[f"maskjpeg{i}"] = np.frombuffer((Path(config["test_mask_path"]) / key(n)).read_bytes(), dtype=np.uint8)
        else:
            out[f"label{i}"] = (np.asarray(__import__("PIL.Image", fromlist=["Image"]).open(
                Path(config["test_mask_path"]) / "3. Hard Exudates" / key(n).replace(".jpg", "_EX.tif")).convert("L")) > 0).astype(np.uint8)
        out[f"pred{i}"] = pred.astype(np.float32)
        out[f"gt{i}"] = gt
        out[f"mask{i}"] = np.asarray(cap["masks"][key(n)])
    return out
