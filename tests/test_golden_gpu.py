"""GPU parity against the committed outputs of the reference itself (tests/golden/) and
pipeline-level parity against the oracle's restatement of the reference loops.  Nothing here
reads /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from eyediseasesegmentation_b200 import _driver as drv, aucpr, kernels as K, ttach_compat as tta  # noqa: E402
from oracle import nets, pipeline, scoring  # noqa: E402
import helpers  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", helpers.NET_GOLDEN_CASES, ids=[c[0] for c in helpers.NET_GOLDEN_CASES])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_cuda_path_reproduces_reference_logits(case, precision, tol):
    key, name, cfg, size, batch = case
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "net_logits.npz"))[key])
    model = helpers.build_product_model(name, cfg).to("cuda")
    model.precision = precision
    out = model(helpers.golden_input(batch, size).cuda()).cpu()
    assert out.shape == ref.shape
    assert (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item() < tol
    bound = 1e-4 if precision == "fp32" else 8e-2
    assert helpers.rel_l2(out - out.mean(), ref - ref.mean()) < bound


def test_scoring_interface_reproduces_reference_outputs(tmp_path):
    golden = json.load(open(os.path.join(GOLDEN, "scoring.json")))
    cfg = {"out_dir": str(tmp_path), "dataset_name": "IDRiD", "lesion_type": "EX"}
    for seed, ref in golden.items():
        items = helpers.synth_scoring_case(int(seed))
        gen = helpers_multigen(items)
        assert aucpr.get_auc(gen, cfg) == pytest.approx(ref["get_auc"], abs=1e-3)         # north_star: AUC-PR within 1e-3
        assert aucpr.get_aucroc(gen, cfg) == pytest.approx(ref["get_aucroc"], abs=1e-3)
        assert list(aucpr.plot_aucpr_curve(gen, "exp", cfg)) == ref["plot_aucpr_curve"]  # integer counts -> same picks
        assert aucpr.plot_aucroc_curve(gen, "exp", cfg) == ref["plot_aucroc_curve"]
        tp, pp, ap, an = aucpr._pooled_counts(gen)
        o = scoring.pr_curve(items)
        assert np.array_equal(tp, o["tp"]) and np.array_equal(pp, o["pp"]) and ap == o["ap"]   # bit-exact counts


def helpers_multigen(items):
    class Re:
        def __iter__(self):
            return iter(items)
    return Re()


def _fundus(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[:h, :w]
    img[(yy - h / 2) ** 2 + (xx - w / 2) ** 2 > (0.48 * min(h, w)) ** 2] = 0      # black outside the disc
    return img


@pytest.mark.parametrize("alias,kind", [("flip_transform", "flip"), ("d4_transform", "d4")])
def test_sliding_window_pipeline_matches_oracle(alias, kind):
    """tta.py:196-213 end to end: tile fetch, TTA net, sigmoid, x2 paste, overwrite order."""
    name, cfg = "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)
    S = 128
    model = helpers.build_product_model(name, cfg)
    sd = model.state_dict()
    image = _fundus(300, 420, 5)
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    want = pipeline.tiled_probability_map(image, lambda t: nets.unetplusplus_forward(sd, t), S, mean, std, kind)
    model = model.to("cuda")
    model.precision = "fp32"
    tfm = getattr(tta.aliases, alias)()
    got = drv.tiled_probability_map(model, tfm, torch.from_numpy(image).cuda(), S, mean, std, tiles_per_batch=3)
    assert got.shape == want.shape
    assert np.abs(got.cpu().numpy() - want).max() < 1e-4
    model.precision = "bf16"
    got16 = drv.tiled_probability_map(model, tfm, torch.from_numpy(image).cuda(), S, mean, std)
    assert np.abs(got16.cpu().numpy() - want).max() < 2e-2


def test_window_larger_than_image_is_rejected():
    model = helpers.build_product_model("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1)).to("cuda")
    img = torch.zeros((200, 200, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError, match="window larger"):
        drv.tiled_probability_map(model, tta.aliases.hflip_transform(), img, 128, [0, 0, 0], [1, 1, 1])


def test_whole_image_resize_matches_cv2():
    prob = torch.rand(1024, 1024, device="cuda")
    full = torch.empty((2848, 4288), dtype=torch.float32, device="cuda")
    K.resize_paste(prob, full, (172, 0, 680, 1024), (0, 0), (2848, 4288))
    want = pipeline.whole_image_probability(prob.cpu().numpy(), (680, 1024), (2848, 4288))
    assert np.abs(full.cpu().numpy() - want).max() < 2e-6


# ---------------------------------------------------------------- full-size properties
def test_full_size_scoring_properties():
    """BASELINE size (2848 x 4288): size-independent properties instead of a CPU re-computation."""
    H, W = 2848, 4288
    g = torch.Generator(device="cuda").manual_seed(0)
    logit = torch.randn(H * W, generator=g, device="cuda") * 3 - 4
    prob = torch.sigmoid(logit)
    gt = (torch.rand(H * W, generator=g, device="cuda") < torch.sigmoid(logit - 1)).to(torch.uint8)
    s = aucpr.score_device(prob.view(H, W), gt.view(H, W))
    assert s.n_pos + s.n_neg == H * W and s.n_pos == int(gt.sum())
    assert np.all(np.diff(s.pp) <= 0) and np.all(np.diff(s.tp) <= 0) and np.all(s.tp <= s.pp)
    assert s.pp[0] == int((prob > 0).sum()) and s.pp[-1] == 0
    for k, t in enumerate(aucpr.thresh_list):                       # exact counts at every threshold
        m = prob.double() > t
        assert s.pp[k] == int(m.sum()) and s.tp[k] == int((m & (gt > 0)).sum())
    perm = torch.randperm(H * W, device="cuda", generator=g)         # order of pixels is irrelevant
    s2 = aucpr.score_device(prob[perm].view(H, W), gt[perm].view(H, W))
    assert s2.ap == pytest.approx(s.ap, abs=1e-12) and np.array_equal(s2.tp, s.tp)
    inv = aucpr.score_device(prob.view(H, W), (1 - gt).view(H, W))   # ROC symmetry under label flip
    assert inv.roc == pytest.approx(1 - s.roc, abs=1e-9)
    sub = slice(0, 1 << 20)                                          # vs sklearn on a 1 Mpx slice
    from sklearn.metrics import average_precision_score
    s3 = aucpr.score_device(prob[sub].view(1024, 1024), gt[sub].view(1024, 1024))
    assert s3.ap == pytest.approx(average_precision_score(gt[sub].cpu().numpy(), prob[sub].cpu().numpy()), abs=1e-3)


def test_tta_merge_linearity_full_tile():
    S, V = 1024, 8
    _, deaug = tta.view_maps(tta.aliases.d4_transform(), S, S)
    a = torch.randn(V, 1, S, S, device="cuda")
    b = torch.randn(V, 1, S, S, device="cuda")
    ma, mb = K.tta_merge(a, deaug, False), K.tta_merge(b, deaug, False)
    mab = K.tta_merge(a + b, deaug, False)
    assert (mab - (ma + mb)).abs().max().item() < 1e-5
    const = K.tta_merge(torch.full((V, 1, S, S), 0.25, device="cuda"), deaug, False)
    assert torch.all(const == 0.25)


def test_confusion_counts_bit_exact():
    """eds_confusion_u8 == numpy on `x > 50` masks: aligned and ragged sizes, several images per launch,
    accumulation into existing counters."""
    import numpy as np
    import torch
    from eyediseasesegmentation_b200 import kernels as K
    rng = np.random.default_rng(5)
    for n_img, n_px in ((1, 16), (3, 1000), (2, 2848 * 4288), (1, 12345)):
        pred = rng.integers(0, 256, size=(n_img, n_px), dtype=np.uint8)
        gt = rng.integers(40, 62, size=(n_img, n_px), dtype=np.uint8)       # dense around the threshold
        got = K.confusion_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda())
        bp, bg = pred > 50, gt > 50
        want = np.stack([(bp & bg).sum(1), bg.sum(1), bp.sum(1)], axis=1)
        assert np.array_equal(got.cpu().numpy(), want)
        again = K.confusion_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), counts=got)
        assert np.array_equal(again.cpu().numpy(), 2 * want)
    # unaligned base pointers take the scalar path
    base_p = torch.from_numpy(rng.integers(0, 256, size=4099, dtype=np.uint8)).cuda()
    base_g = torch.from_numpy(rng.integers(0, 256, size=4099, dtype=np.uint8)).cuda()
    got = K.confusion_counts(base_p[3:].view(1, -1), base_g[3:].view(1, -1), 100, 7)
    bp, bg = base_p[3:].cpu().numpy() > 100, base_g[3:].cpu().numpy() > 7
    assert got.cpu().numpy().tolist() == [[int((bp & bg).sum()), int(bg.sum()), int(bp.sum())]]


def test_export_result_writes_the_reference_csvs(tmp_path):
    """stat_result.export_result / stat_result_vessel.export_result on the GPU reproduce the CSV files of
    the reference (golden fixture) for the seeded mask set."""
    import json
    import os
    import helpers
    from eyediseasesegmentation_b200 import stat_result, stat_result_vessel
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stat_result.json")))
    lesion_cfg, vessel_cfg = helpers.make_stat_case(tmp_path, seed=0)
    stat_result.export_result("EX/exp", lesion_cfg)
    stat_result_vessel.export_result("vexp", vessel_cfg)
    helpers.assert_stat_csvs_equal(
        helpers.read_stat_csvs(tmp_path / "out" / "IDRiD" / "result_assessment" / "EX" / "exp"), golden["lesion"])
    helpers.assert_stat_csvs_equal(
        helpers.read_stat_csvs(tmp_path / "out" / "DRIVE" / "result_assessment" / "vexp"), golden["vessel"])


def test_aucpr_callback_matches_reference_metric():
    """util/aucpr_cb.py: batches arrive one by one (logits [B,1,H,W], float targets), the loader metric is the
    trapezoid PR-AUC over all their pixels; restarting the loader resets the state."""
    from types import SimpleNamespace
    from eyediseasesegmentation_b200.aucpr_cb import AucPRMetricCallback
    from oracle import scoring
    g = torch.Generator().manual_seed(31)
    cb = AucPRMetricCallback()
    runner = SimpleNamespace(output={}, input={}, loader_metrics={})
    for epoch in range(2):
        cb.on_loader_start(runner)
        y_trues, y_preds = [], []
        for b in range(3):
            targets = (torch.rand(2 + b, 1, 96, 80, generator=g) < 0.05).float()
            logits = torch.randn(2 + b, 1, 96, 80, generator=g) * 2 + (targets * 2 - 1) * (1.0 + epoch)
            runner.output = {"logits": logits.cuda()}
            runner.input = {"targets": targets.cuda()}
            cb.on_batch_end(runner)
            y_trues.extend(targets.numpy())
            y_preds.extend(torch.sigmoid(logits).numpy())
        cb.on_loader_end(runner)
        want = scoring.callback_pr_auc(y_trues, y_preds)
        assert abs(runner.loader_metrics["auc_pr"] - want) < 1e-4, (epoch, runner.loader_metrics, want)


def test_gaussian_tile_blend_driver_matches_oracle_restatement(monkeypatch):
    """_driver.tiled_probability_map(blend="gaussian") (opt-in; config["tile_blend"] / EDS_TILE_BLEND) against
    oracle/pipeline.py's numpy restatement with the same network, fp32 mode: 1e-4.  Overwrite stays the default."""
    from eyediseasesegmentation_b200 import _driver as drv
    assert drv.tile_blend_mode({}) == "overwrite" and drv.tile_blend_mode({"tile_blend": "gaussian"}) == "gaussian"
    with pytest.raises(ValueError):
        drv.tile_blend_mode({"tile_blend": "pyramid"})
    name, cfg = "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)
    model = helpers.build_product_model(name, cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["segmentation_head.0.weight"] *= 6.0
    model.load_state_dict(sd)
    S = 64
    rng = np.random.default_rng(4)
    image = rng.integers(0, 256, size=(200, 262, 3), dtype=np.uint8)
    mean, std = pipeline.DATASET_STATS["IDRiD"]
    want = pipeline.tiled_probability_map(image, lambda t: nets.unetplusplus_forward(sd, t), S, mean, std, "d4",
                                          blend="gaussian")
    plain = pipeline.tiled_probability_map(image, lambda t: nets.unetplusplus_forward(sd, t), S, mean, std, "d4")
    model = model.to("cuda")
    model.precision = "fp32"
    got = drv.tiled_probability_map(model, tta.aliases.d4_transform(), torch.from_numpy(image).cuda(), S, mean, std,
                                    blend="gaussian").cpu().numpy()
    assert np.abs(got - want).max() < 1e-4
    assert np.abs(want - plain).max() > 1e-3                 # the mode does something


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-3)])
def test_ensemble_predict_reproduces_the_reference_ensemble_script(tmp_path, monkeypatch, precision, tol):
    """tests/golden/ensemble.npz holds what the reference's own ensemble.py (run unmodified, make_golden.py) produced
    on helpers.make_ensemble_case: the product's ensemble.predict, started in the same folder layout (the dataset
    paths ensemble.py:65-66 hard-codes, relative to the working directory), must return its AUC-PR and write its
    masks."""
    from PIL import Image
    from eyediseasesegmentation_b200 import ensemble as eds_ensemble
    g = np.load(os.path.join(GOLDEN, "ensemble.npz"))
    monkeypatch.setenv("EDS_PRECISION", precision)
    config, logdirs = helpers.make_ensemble_case(tmp_path, gts=g["gts"])
    monkeypatch.chdir(tmp_path)
    got_auc = eds_ensemble.predict(config, [os.path.relpath(d, tmp_path) for d in logdirs], "ensemble_1")
    assert abs(got_auc - float(g["auc"])) < tol                 # BASELINE: AUC-PR within 1e-3 (bf16)
    written = tmp_path / "outputs" / "IDRiD" / "tta" / "EX" / "ensemble_1"
    assert sorted(p.name for p in written.iterdir()) == [str(n) for n in g["names"]]
    t1 = float(g["thresholds"][0])
    for name, want, pred in zip(g["names"], g["masks"], g["preds"]):
        got = np.asarray(Image.open(written / str(name)).convert("L")) > 127
        undecided = np.abs(pred - t1) < (1e-3 if precision == "fp32" else 2e-2)
        assert got.shape == want.shape
        # pixels a rounding difference cannot flip; the masks travel through the reference's JPEG writer
        assert np.mean(got[~undecided] != want.astype(bool)[~undecided]) < 2e-3, name
