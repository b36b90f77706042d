"""CPU tests (no GPU): the oracle against the reference's own outputs.

  - committed golden fixtures (tests/golden/, produced by the reference code itself via
    tests/golden/make_golden.py) pin oracle/nets.py, oracle/scoring.py and oracle/pipeline.py
    everywhere, including the GPU box where /root/reference does not exist;
  - when /root/reference is present (build container) the oracle is additionally compared
    with the reference modules directly, bit for bit.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nets, pipeline, ref_loader, scoring
import helpers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HAS_REF = ref_loader.available()


def _golden_logits():
    return np.load(os.path.join(GOLDEN, "net_logits.npz"))


@pytest.mark.parametrize("case", helpers.NET_GOLDEN_CASES, ids=[c[0] for c in helpers.NET_GOLDEN_CASES])
def test_oracle_nets_reproduce_reference_logits(case):
    key, name, cfg, size, batch = case
    model = helpers.build_product_model(name, cfg)
    out = helpers.oracle_forward(name, cfg, model.state_dict(), helpers.golden_input(batch, size))
    ref = torch.from_numpy(_golden_logits()[key])
    assert out.shape == ref.shape
    # same ops in the same order on the same CPU build: exact in the build container; allow a few
    # ulp for a different host CPU's conv algorithm choice
    assert (out - ref).abs().max().item() < 2e-5


def test_state_dict_layout_matches_reference():
    layout = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))
    for key, name, cfg, _, _ in helpers.NET_GOLDEN_CASES:
        mine = {k: list(v.shape) for k, v in helpers.build_product_model(name, cfg).state_dict().items()}
        assert mine == layout[key], key


def test_shared_axial_block_aliases_storage():
    sd = helpers.build_product_model("unetplusplusstar", helpers.star_cfg(8)).state_dict()
    assert len(sd) == 869 and len({v.data_ptr() for v in sd.values()}) == 821
    a, b = sd["encoder.layer4.1.in_conv1x1.0.weight"], sd["encoder.layer4.2.in_conv1x1.0.weight"]
    assert a.data_ptr() == b.data_ptr()


@pytest.mark.parametrize("bd,count", [(8, 60167435), (16, 60179723), (19, 60184331), (32, 60204299)])
def test_parameter_counts(bd, count):
    m = helpers.build_product_model("unetplusplusstar", helpers.star_cfg(bd))
    assert sum(p.numel() for p in m.parameters()) == count          # SURVEY.md 8c known answers


def test_parameter_counts_baselines():
    counts = {"upp_se50_scse": 53658171, "upp_r34": 26080340, "unet_r34": 24436369}
    for key, name, cfg, _, _ in helpers.NET_GOLDEN_CASES[1:]:
        m = helpers.build_product_model(name, cfg)
        assert sum(p.numel() for p in m.parameters()) == counts[key]


def test_oracle_scoring_matches_reference_golden():
    golden = json.load(open(os.path.join(GOLDEN, "scoring.json")))
    for seed, ref in golden.items():
        items = helpers.synth_scoring_case(int(seed))
        assert scoring.get_auc(items) == pytest.approx(ref["get_auc"], abs=1e-12)
        assert scoring.get_aucroc(items) == pytest.approx(ref["get_aucroc"], abs=1e-12)
        assert list(scoring.pr_curve(items)["thresholds"]) == ref["plot_aucpr_curve"]
        assert scoring.roc_curve(items)["threshold"] == ref["plot_aucroc_curve"]


def test_make_grid_matches_reference_golden():
    from eyediseasesegmentation_b200.util import make_grid
    for case in json.load(open(os.path.join(GOLDEN, "make_grid.json"))):
        want = np.array(case["grid"], dtype=np.int64)
        for fn in (pipeline.make_grid, make_grid):
            got = fn(tuple(case["shape"]), window=case["window"], min_overlap=case["min_overlap"])
            assert got.dtype == np.int64 and np.array_equal(got, want), (fn.__module__, case["shape"])


def test_registry_and_preprocessing_match_reference_golden():
    from eyediseasesegmentation_b200 import archs
    g = json.load(open(os.path.join(GOLDEN, "registry_preprocessing.json")))
    assert archs.list_models() == list(dict.fromkeys(g["registry"]))
    sample = np.arange(0, 256, 5, dtype=np.uint8).reshape(-1, 1, 1).repeat(3, axis=2)
    for ds, ref in g["preprocessing"].items():
        fn, mean, std = archs.get_preprocessing_fn(None if ds == "None" else ds, False)
        assert mean == ref["mean"] and std == ref["std"]
        assert np.array_equal(fn(sample).astype(np.float32).reshape(-1), np.array(ref["out"], dtype=np.float32))
    with pytest.raises(KeyError):
        archs.get_model("no_such_model", {}, training=False)
    with pytest.raises(NotImplementedError):
        archs.get_model("hrnet18", {}, training=False)


def test_get_model_applies_inference_overrides():
    from eyediseasesegmentation_b200 import archs
    params = dict(encoder_name="resnet34", encoder_weights="imagenet", classes=1, deep_supervision=True)
    archs.get_model("unetplusplus_deepsup", params, training=False)
    assert params["encoder_weights"] is None and params["deep_supervision"] is False   # archs/__init__.py:111-119


# ------------------------------------------------------------ direct, when the reference is here
@pytest.mark.skipif(not HAS_REF, reason="/root/reference is only present in the build container")
def test_oracle_nets_bit_exact_against_reference_modules():
    ref = ref_loader.load()
    x = helpers.golden_input(1, 256)
    torch.manual_seed(3)
    m = ref.unetplusplusstar.UnetPlusPlusStar(**helpers.star_cfg(8)).eval()
    helpers.randomize_bn(m, 5)
    with torch.no_grad():
        assert torch.equal(m(x), nets.unetplusplusstar_forward(m.state_dict(), x, 8))
        m2 = ref.deep_supunetplusplus.UnetPlusPlus(encoder_name="se_resnet50", encoder_weights=None, classes=1,
                                                   decoder_attention_type="scse").eval()
        helpers.randomize_bn(m2, 6)
        assert torch.equal(m2(x), nets.unetplusplus_forward(m2.state_dict(), x))
        m3 = ref.smp.Unet(encoder_name="resnet34", encoder_weights=None, classes=1).eval()
        helpers.randomize_bn(m3, 7)
        assert torch.equal(m3(x), nets.unet_forward(m3.state_dict(), x))


@pytest.mark.skipif(not HAS_REF, reason="/root/reference is only present in the build container")
def test_survey_known_answers_of_the_reference_network():
    """SURVEY.md 8c known answers: the reference's own constructor under torch.manual_seed(1999) (pipeline.py:36),
    proposed network, base_dim 8, x = randn(1,3,256,256) from seed 7 -> logit statistics, two probe pixels, and
    the statistics of the D4-TTA mean (the oracle's restatement of ttach on the reference's weights)."""
    ref = ref_loader.load()
    torch.manual_seed(1999)
    m = ref.unetplusplusstar.UnetPlusPlusStar(**helpers.star_cfg(8)).eval()
    x = helpers.golden_input(1, 256, seed=7)
    sd = m.state_dict()
    with torch.no_grad():
        y = nets.unetplusplusstar_forward(sd, x, 8)
        assert torch.equal(y, m(x))
        t = nets.tta_mean_logits(lambda z: nets.unetplusplusstar_forward(sd, z, 8), x, "d4")
    assert y.mean().item() == pytest.approx(-0.551582, abs=2e-6) and y.std().item() == pytest.approx(0.139031, abs=2e-6)
    assert y[0, 0, 0, 0].item() == pytest.approx(0.006472, abs=2e-6)
    assert y[0, 0, 128, 128].item() == pytest.approx(-0.619326, abs=2e-6)
    assert t.mean().item() == pytest.approx(-0.553268, abs=2e-6) and t.std().item() == pytest.approx(0.089578, abs=2e-6)


@pytest.mark.skipif(not HAS_REF, reason="/root/reference is only present in the build container")
def test_oracle_scoring_against_reference_aucpr(tmp_path):
    ref = ref_loader.load()
    cfg = {"out_dir": str(tmp_path), "dataset_name": "IDRiD", "lesion_type": "EX"}
    items = helpers.synth_scoring_case(11, shape=(64, 80), n_images=4)
    assert scoring.get_auc(items) == ref.aucpr.get_auc(items, cfg)
    assert scoring.get_aucroc(items) == ref.aucpr.get_aucroc(items, cfg)
    assert scoring.pr_curve(items)["thresholds"] == tuple(ref.aucpr.plot_aucpr_curve(items, "t", cfg))
    assert scoring.roc_curve(items)["threshold"] == ref.aucpr.plot_aucroc_curve(items, "t", cfg)
    with pytest.raises(ZeroDivisionError):
        ref.aucpr.get_auc([(items[0][0], np.zeros_like(items[0][1]), "empty")], cfg)
    with pytest.raises(ZeroDivisionError):
        scoring.get_auc([(items[0][0], np.zeros_like(items[0][1]), "empty")])


def test_histogram_key_keeps_ap_within_budget():
    """DESIGN.md 'score key': AP on key-quantised scores vs AP on raw fp32 scores."""
    from sklearn.metrics import average_precision_score
    worst = 0.0
    for seed in range(4):
        for pred, gt, _ in helpers.synth_scoring_case(seed, shape=(128, 160), n_images=3)[:2]:
            a = average_precision_score(gt.reshape(-1), pred.reshape(-1))
            b = average_precision_score(gt.reshape(-1), scoring.score_key(pred).reshape(-1))
            worst = max(worst, abs(a - b))
    # 20 K-pixel images are the hard case (one tie is worth 1/n_pos); the budget is 1e-3
    assert worst < 5e-4, worst


def test_histogram_key_resolves_saturated_scores():
    """Confident (saturated) sigmoid outputs: the key is symmetric about 1/2, so scores within 1e-3 of 1 keep
    their order.  AP / ROC-AUC / the callback's trapezoid PR-AUC of key-quantised scores stay within 1e-4."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    rng = np.random.default_rng(21)
    for prevalence, sharp, sep in [(0.02, 4.0, 1.5), (0.01, 3.0, 4.0), (0.01, 6.0, 6.0), (0.05, 8.0, 3.0)]:
        gt = (rng.random(300_000) < prevalence).astype(np.uint8)
        logit = rng.normal(size=gt.size) * sharp + (gt.astype(np.float64) * 2 - 1) * sep
        pred = (1.0 / (1.0 + np.exp(-logit))).astype(np.float32)
        key = scoring.score_key(pred)
        order = np.argsort(pred, kind="stable")
        assert np.all(np.diff(key[order]) >= 0)                       # monotone
        assert abs(average_precision_score(gt, pred) - average_precision_score(gt, key)) < 1e-4
        assert abs(roc_auc_score(gt, pred) - roc_auc_score(gt, key)) < 1e-4
        assert abs(scoring.callback_pr_auc([gt], [pred]) - scoring.callback_pr_auc([gt], [key.astype(np.float64)])) < 1e-4
    edge = np.array([0.0, 2.0 ** -25, 2.0 ** -24, 0.25, np.nextafter(np.float32(0.5), np.float32(0)), 0.5,
                     np.nextafter(np.float32(0.5), np.float32(1)), 1 - 2.0 ** -24, 1.0], dtype=np.float32)
    k = scoring.score_key(edge)
    assert k[0] == 0 and k[1] == 0 and k[2] == 1 and k[-1] == 2 * (23 * 512 + 2) - 1 and np.all(np.diff(k) >= 0)
    assert k[5] - k[4] == 2 and k[6] - k[5] == 1          # q = 1/2 has its own bin; the low half's top bin stays empty


def test_stat_result_oracle_matches_reference_csvs(tmp_path):
    """oracle/stat_result.py reproduces the CSV files the reference's own export_result wrote for the
    seeded mask set (tests/golden/stat_result.json, generated by make_golden.py); in the build
    container the reference itself is run again next to it."""
    import json
    import helpers
    from oracle import stat_result as osr
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stat_result.json")))
    lesion_cfg, vessel_cfg = helpers.make_stat_case(tmp_path, seed=0)
    gt_dir = str(lesion_cfg["test_mask_path"] / "3. Hard Exudates")
    osr.export_result(gt_dir, lesion_cfg["out_dir"] + "/IDRiD/tta/EX/exp", osr.lesion_pred_name("IDRiD", "EX"),
                      str(tmp_path / "o_lesion"))
    osr.export_result(str(vessel_cfg["test_mask_path"]), vessel_cfg["out_dir"] + "/DRIVE/tta/vexp", lambda n: n,
                      str(tmp_path / "o_vessel"))
    helpers.assert_stat_csvs_equal(helpers.read_stat_csvs(tmp_path / "o_lesion"), golden["lesion"])
    helpers.assert_stat_csvs_equal(helpers.read_stat_csvs(tmp_path / "o_vessel"), golden["vessel"])
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load()
        ref.stat_result.export_result("EX/exp", lesion_cfg)
        helpers.assert_stat_csvs_equal(
            helpers.read_stat_csvs(tmp_path / "out" / "IDRiD" / "result_assessment" / "EX" / "exp"), golden["lesion"])


# ------------------------------------------------------------------ vessel padding (pad_img.py:8-38)
def _pad_golden():
    return np.load(os.path.join(GOLDEN, "pad_img.npz"))


def test_pad_img_oracle_and_product_match_reference_golden(tmp_path):
    """The reference's own pad() wrote tests/golden/pad_img.npz (file to file).  Both the oracle restatement (array
    level, same cv2 calls) and the product mirror (file to file, same signature) must reproduce it byte for byte:
    centre pad with top = dh // 2 / left = dw // 2 (odd deltas put the extra line at the bottom / right) and
    the `> 127 -> 255` label threshold."""
    import cv2
    from oracle import pad as opad
    from eyediseasesegmentation_b200 import pad_img
    golden = _pad_golden()
    img_in, lab_in, size = helpers.make_pad_case(tmp_path, seed=0)
    pad_img.pad(str(img_in), str(tmp_path / "o_img"), desired_size=size)
    pad_img.pad(str(lab_in), str(tmp_path / "o_lab"), desired_size=size, is_mask=True)
    got = {"img/" + k: v for k, v in helpers.read_pad_outputs(tmp_path / "o_img").items()}
    got.update({"lab/" + k: v for k, v in helpers.read_pad_outputs(tmp_path / "o_lab").items()})
    assert sorted(got) == sorted(golden.files)
    for key in golden.files:
        assert got[key].shape == golden[key].shape and np.array_equal(got[key], golden[key]), key
        folder, name = key.split("/")
        src = cv2.imread(str((img_in if folder == "img" else lab_in) / name), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(opad.pad_array(src, size, folder == "lab"), golden[key]), key
        assert np.array_equal(pad_img.pad_array(src, size, folder == "lab"), golden[key]), key


def test_pad_img_vessel_geometries():
    """DRIVE 584x565 -> 608^2 pads 12/12 rows and 21/22 columns; CHASEDB1 960x999 -> 1024^2 pads
    32/32 and 12/13."""
    from eyediseasesegmentation_b200 import pad_img
    assert pad_img.pad_geometry((584, 565), 608) == (12, 12, 21, 22)
    assert pad_img.pad_geometry((960, 999), 1024) == (32, 32, 12, 13)
    with pytest.raises(ValueError):
        pad_img.pad_geometry((700, 565), 608)
    already = np.arange(16, dtype=np.uint8).reshape(4, 4)
    assert np.array_equal(pad_img.pad_array(already, 4), already)


@pytest.mark.skipif(not HAS_REF, reason="reference tree not present")
def test_pad_img_matches_reference_function(tmp_path):
    from eyediseasesegmentation_b200 import pad_img
    refpad = ref_loader.load_pad_img()
    img_in, lab_in, size = helpers.make_pad_case(tmp_path, seed=3)
    for folder, is_mask in ((img_in, False), (lab_in, True)):
        refpad.pad(str(folder), str(tmp_path / "ref"), desired_size=size + 8, is_mask=is_mask)
        pad_img.pad(str(folder), str(tmp_path / "got"), desired_size=size + 8, is_mask=is_mask)
        want, got = helpers.read_pad_outputs(tmp_path / "ref"), helpers.read_pad_outputs(tmp_path / "got")
        assert sorted(want) == sorted(got)
        for k in want:
            assert np.array_equal(want[k], got[k]), (k, is_mask)


# ------------------------------------------------------------------ ensemble.py
def _golden_ensemble():
    return np.load(os.path.join(GOLDEN, "ensemble.npz"))


def test_ensemble_oracle_matches_the_reference_ensemble_script():
    """tests/golden/ensemble.npz was written by the reference's OWN ensemble.py (get_best_model + predict, run
    unmodified by make_golden.py through ref_loader.load_ensemble).  oracle.pipeline.ensemble_probability + the
    oracle scoring must reproduce its per-image probability maps, AUC-PR, threshold and masks."""
    g = _golden_ensemble()
    nets_list = helpers.ensemble_oracle_nets(helpers.ensemble_state_dicts())
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    items = []
    for img, gt, want, name in zip(g["images"], g["gts"], g["preds"], g["names"]):
        x = torch.from_numpy(pipeline.preprocess(img, mean, std).transpose(2, 0, 1)).float()[None]
        pred = pipeline.ensemble_probability(nets_list, x)
        assert pred.dtype == np.float32 and pred.shape == want.shape
        assert np.abs(pred - want).max() < 2e-6, name          # exact here; a few ulp for another host CPU
        items.append((pred, gt, str(name)))
    assert abs(scoring.get_auc(items) - float(g["auc"])) < 1e-6
    t1 = scoring.pr_curve(items)["thresholds"][0]
    assert t1 == float(g["thresholds"][0])
    for (pred, _, name), want_mask in zip(items, g["masks"]):
        decided = np.abs(pred - t1) > 1e-5                      # pixels a last-ulp difference cannot flip
        assert np.array_equal((pred > t1)[decided], want_mask.astype(bool)[decided]), name


@pytest.mark.skipif(not HAS_REF, reason="needs /root/reference (build container)")
def test_reference_ensemble_script_reproduces_its_fixture(tmp_path):
    """The committed fixture is what the reference's ensemble.py produces today (regeneration check)."""
    g = _golden_ensemble()
    out = helpers.run_reference_ensemble(tmp_path, gts=g["gts"])
    assert list(out["names"]) == list(g["names"])
    assert np.array_equal(out["images"], g["images"]) and np.array_equal(out["gts"], g["gts"])
    assert np.abs(out["preds"] - g["preds"]).max() < 2e-6
    assert np.array_equal(out["masks"], g["masks"])
    assert abs(float(out["auc"]) - float(g["auc"])) < 1e-9 and list(out["thresholds"]) == list(g["thresholds"])


# ------------------------------------------------------------------ tta.py drivers
def _golden_tta(case):
    g = np.load(os.path.join(GOLDEN, f"tta_{case}.npz"))
    return g, len(g["names"])


def _decode(jpeg_bytes):
    import io
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(bytes(jpeg_bytes))).convert("RGB")).astype("uint8")


def _check_scores_and_masks(g, items, mask_dtype):
    assert abs(scoring.get_auc(items) - float(g["auc"])) < 1e-6
    th = scoring.pr_curve(items)["thresholds"]
    assert list(th) == list(g["thresholds"])
    for i, (pred, _, name) in enumerate(items):
        want = g[f"mask{i}"]
        assert want.dtype == mask_dtype                         # tta.py:226 writes float32, tta.py:138 uint8
        decided = np.abs(pred - th[2]) > 1e-5                   # pixels a last-ulp difference cannot flip
        assert np.array_equal((pred > th[2])[decided], want.astype(bool)[decided]), name


def test_tiled_pipeline_oracle_matches_the_reference_tta_patches():
    """tests/golden/tta_patches.npz was written by the reference's OWN tta.tta_patches (tta.py:150-238, run
    unmodified by make_golden.py through ref_loader.load_tta: windowed read, A.Resize, preprocessing, D4 TTA of the
    reference's UNet++ module, sigmoid, cv2 x2 resize, overwrite paste, get_auc, plot_aucpr_curve, the third
    threshold, the float32 masks).  oracle.pipeline.tiled_probability_map + oracle.scoring must reproduce its
    probability maps, scores, thresholds and masks from the same JPEG bytes."""
    g, n = _golden_tta("patches")
    _, _, S, alias, shapes = helpers.TTA_CASES["patches"]
    net = helpers.tta_case_oracle_net("patches", helpers.tta_case_state_dict("patches"))
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    items = []
    for i in range(n):
        image = _decode(g[f"jpeg{i}"])
        assert image.shape[:2] == shapes[i]
        pred = pipeline.tiled_probability_map(image, net, S, mean, std, alias)
        assert np.abs(pred - g[f"pred{i}"]).max() < 2e-6, i     # exact here; a few ulp for another host CPU
        assert np.array_equal(g[f"gt{i}"], g[f"label{i}"])      # tta.py:192-194: labels scored as stored (> 0)
        items.append((pred, g[f"gt{i}"], str(g["names"][i])))
    assert items[-1][1].sum() == 0                              # the image aucpr.py:22 skips
    _check_scores_and_masks(g, items, np.float32)


def test_whole_image_oracle_matches_the_reference_test_tta():
    """tests/golden/tta_whole.npz: the reference's OWN tta.test_tta (tta.py:56-148) on two 89 x 134 images through
    its TestSegmentation / NormalTransform (LongestMaxSize(1024) + centred pad), smp.Unet(resnet34) under hflip TTA,
    sigmoid, centre crop + cv2 resize of prediction AND mask back to the original size, scoring, uint8 masks."""
    import cv2
    g, n = _golden_tta("whole")
    _, _, S, alias, shapes = helpers.TTA_CASES["whole"]
    net = helpers.tta_case_oracle_net("whole", helpers.tta_case_state_dict("whole"))
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    items = []
    for i in range(n):
        image = _decode(g[f"jpeg{i}"])
        H0, W0 = image.shape[:2]
        scale = S / max(H0, W0)
        nh, nw = int(round(H0 * scale)), int(round(W0 * scale))
        top, left = int((S - nh) / 2.0), int((S - nw) / 2.0)
        padded = np.zeros((S, S, 3), dtype=np.uint8)
        padded[top:top + nh, left:left + nw] = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_LINEAR)
        x = torch.from_numpy(pipeline.preprocess(padded, mean, std).transpose(2, 0, 1)).float()[None]
        with torch.no_grad():
            prob = torch.sigmoid(nets.tta_mean_logits(net, x, alias)[0, 0]).numpy()
        pred = pipeline.whole_image_probability(prob, (nh, nw), (H0, W0))
        assert np.abs(pred - g[f"pred{i}"]).max() < 2e-6, i
        mp = np.zeros((S, S), dtype=np.uint8)                   # lesion_dataset.py:126-128 + the same transform
        mp[top:top + nh, left:left + nw] = cv2.resize(g[f"label{i}"], (nw, nh), interpolation=cv2.INTER_NEAREST)
        gt = pipeline.whole_image_probability(mp, (nh, nw), (H0, W0))
        assert np.array_equal(gt, g[f"gt{i}"])
        items.append((pred, gt, str(g["names"][i])))
    _check_scores_and_masks(g, items, np.uint8)


@pytest.mark.skipif(not HAS_REF, reason="needs /root/reference (build container)")
def test_reference_tta_patches_reproduces_its_fixture(tmp_path):
    """The committed sliding-window fixture is what the reference's tta.py produces today (regeneration check)."""
    g, n = _golden_tta("patches")
    out = helpers.run_reference_tta(tmp_path, "patches", jpegs=[g[f"jpeg{i}"] for i in range(n)],
                                    gts=[g[f"label{i}"] for i in range(n)])
    assert list(out["names"]) == list(g["names"]) and list(out["thresholds"]) == list(g["thresholds"])
    assert abs(float(out["auc"]) - float(g["auc"])) < 1e-9
    for i in range(n):
        assert np.abs(out[f"pred{i}"] - g[f"pred{i}"]).max() < 2e-6
        assert np.array_equal(out[f"mask{i}"], g[f"mask{i}"]) and np.array_equal(out[f"gt{i}"], g[f"gt{i}"])


def test_vessel_oracle_matches_the_reference_tta_vessel():
    """tests/golden/tta_vessel.npz: the reference's OWN tta_vessel.test_tta (tta_vessel.py:55-136) on two pre-padded
    256 x 256 images -- its TestSegmentation (labels read with > 50), dataset-independent statistics
    (get_preprocessing_fn(None)), the reference's proposed network (base_dim 8) under D4 TTA, sigmoid, get_aucroc,
    plot_aucroc_curve's threshold, uint8 masks.  The oracle must reproduce all of it from the same JPEG bytes."""
    g, n = _golden_tta("vessel")
    _, _, S, alias, _ = helpers.TTA_CASES["vessel"]
    net = helpers.tta_case_oracle_net("vessel", helpers.tta_case_state_dict("vessel"))
    mean, std = pipeline.DATASET_STATS["IDRiD"]                 # tta_vessel.py:73 passes dataset_name=None
    items = []
    for i in range(n):
        image = _decode(g[f"jpeg{i}"])
        gt = (_decode(g[f"maskjpeg{i}"])[..., 0] > 50).astype(np.uint8)
        assert image.shape[:2] == (S, S) and np.array_equal(gt, g[f"gt{i}"])
        x = torch.from_numpy(pipeline.preprocess(image, mean, std).transpose(2, 0, 1)).float()[None]
        with torch.no_grad():
            pred = torch.sigmoid(nets.tta_mean_logits(net, x, alias)[0, 0]).numpy()
        assert np.abs(pred - g[f"pred{i}"]).max() < 2e-6, i
        items.append((pred, gt, str(g["names"][i])))
    assert abs(scoring.get_aucroc(items) - float(g["auc"])) < 1e-6
    t = scoring.roc_curve(items)["threshold"]
    assert [t] == list(g["thresholds"])
    for i, (pred, _, name) in enumerate(items):
        want = g[f"mask{i}"]
        assert want.dtype == np.uint8
        decided = np.abs(pred - t) > 1e-5
        assert np.array_equal((pred > t)[decided], want.astype(bool)[decided]), name


def test_vessel_tiled_oracle_matches_the_reference_tta_vessel_patches():
    """tests/golden/tta_vpatches.npz: the reference's OWN tta_vessel.tta_patches (tta_vessel.py:138-229) on two
    unpadded images -- PIL read, make_grid windows of 2S, A.Resize, DRIVE statistics, D4 TTA, sigmoid, cv2 x2 resize,
    overwrite paste, labels read with > 50, get_aucroc, plot_aucroc_curve's threshold, float32 masks."""
    g, n = _golden_tta("vpatches")
    _, _, S, alias, shapes = helpers.TTA_CASES["vpatches"]
    net = helpers.tta_case_oracle_net("vpatches", helpers.tta_case_state_dict("vpatches"))
    mean, std = pipeline.DATASET_STATS["DRIVE"]
    items = []
    for i in range(n):
        image = _decode(g[f"jpeg{i}"])
        gt = (_decode(g[f"maskjpeg{i}"])[..., 0] > 50).astype(np.uint8)
        assert image.shape[:2] == shapes[i] and np.array_equal(gt, g[f"gt{i}"])
        pred = pipeline.tiled_probability_map(image, net, S, mean, std, alias)
        assert np.abs(pred - g[f"pred{i}"]).max() < 2e-6, i
        items.append((pred, gt, str(g["names"][i])))
    assert abs(scoring.get_aucroc(items) - float(g["auc"])) < 1e-6
    t = scoring.roc_curve(items)["threshold"]
    assert [t] == list(g["thresholds"])
    for i, (pred, _, name) in enumerate(items):
        want = g[f"mask{i}"]
        assert want.dtype == np.float32                         # tta_vessel.py:217
        decided = np.abs(pred - t) > 1e-5
        assert np.array_equal((pred > t)[decided], want.astype(bool)[decided]), name
