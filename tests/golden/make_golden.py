"""Generates the committed fixtures of tests/golden/ by running the REFERENCE'S OWN code
(loaded by oracle/ref_loader.py) in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures travel to the GPU box (which has no /root/reference) and pin both the oracle
(tests/test_oracle.py, CPU) and the CUDA path (tests/test_golden_gpu.py).  Inputs are
re-created from seeds on both sides: the weights come from the product's deterministic
initialiser and are loaded into the reference modules with load_state_dict(strict=True), which
also pins the state_dict key layout.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import helpers  # noqa: E402


def reference_module(ref, name, cfg):
    return ref_loader.build_reference_module(ref, name, cfg)


def pad_fixture_only():
    import tempfile
    refpad = ref_loader.load_pad_img()
    with tempfile.TemporaryDirectory() as tmp:
        img_in, lab_in, size = helpers.make_pad_case(tmp, seed=0)
        refpad.pad(str(img_in), os.path.join(tmp, "o_img"), desired_size=size)
        refpad.pad(str(lab_in), os.path.join(tmp, "o_lab"), desired_size=size, is_mask=True)
        arrays = {"img/" + k: v for k, v in helpers.read_pad_outputs(os.path.join(tmp, "o_img")).items()}
        arrays.update({"lab/" + k: v for k, v in helpers.read_pad_outputs(os.path.join(tmp, "o_lab")).items()})
    np.savez_compressed(os.path.join(HERE, "pad_img.npz"), **arrays)
    print("pad_img.npz written:", sorted(arrays))


def ensemble_fixture_only():
    """The reference's own top-level ensemble.py (predict + get_best_model, unmodified; see
    oracle/ref_loader.load_ensemble for the third-party restatements and the four adapters) on two member models and
    three seeded 64 x 64 images: per-image ensemble probability maps, AUC-PR, thresholds and the written masks."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        first = helpers.run_reference_ensemble(tmp)               # random labels: only to learn the probabilities
    # labels for the committed case: the top 30 % of each probability map with 5 % of the 8 x 8 blocks flipped, so
    # that precision / recall move over the threshold list and the written masks have content
    rng = np.random.default_rng(3)
    gts = []
    for pred in first["preds"]:
        flips = np.kron(rng.random((pred.shape[0] // 8, pred.shape[1] // 8)) < 0.05, np.ones((8, 8))).astype(bool)
        gts.append(((pred > np.quantile(pred, 0.7)) ^ flips).astype(np.uint8))
    with tempfile.TemporaryDirectory() as tmp:
        out = helpers.run_reference_ensemble(tmp, gts=np.stack(gts))
    assert np.array_equal(out["gts"], np.stack(gts)) and np.array_equal(out["preds"], first["preds"])
    np.savez_compressed(os.path.join(HERE, "ensemble.npz"), **out)
    print("ensemble.npz written:", {k: getattr(v, "shape", v) for k, v in out.items()}, "auc", float(out["auc"]),
          "thresholds", out["thresholds"].tolist())


def tta_fixtures_only():
    """The reference's own src/main/tta.py -- tta_patches (sliding window, D4) and test_tta (whole image, hflip) -- and
    src/main/tta_vessel.py -- test_tta (pre-padded squares, proposed network, D4, ROC scoring) and tta_patches (sliding
    window over unpadded images, DRIVE statistics, ROC scoring) -- run unmodified
    (oracle/ref_loader.load_tta) from JPEG / TIFF files to scores and masks.  Two passes: the first, on
    random labels, only learns the probability maps; the committed case uses labels derived from them (top 30 % of
    each map with 5 % of the 8 x 8 blocks flipped; one image without positives for aucpr.py:22) so that the PR
    curve, the chosen threshold and the written masks are not degenerate."""
    import tempfile
    for case in ("patches", "whole", "vessel", "vpatches"):
        with tempfile.TemporaryDirectory() as tmp:
            first = helpers.run_reference_tta(tmp, case)
        n = len(first["names"])
        rng = np.random.default_rng(17)
        gts = []
        for i in range(n):
            pred = first[f"pred{i}"]
            top = np.zeros(pred.size, dtype=bool)
            top[np.argsort(pred.ravel(), kind="stable")[-int(0.3 * pred.size):]] = True
            h, w = pred.shape
            flips = np.kron(rng.random((h // 8 + 1, w // 8 + 1)) < 0.05, np.ones((8, 8)))[:h, :w].astype(bool)
            gts.append((top.reshape(pred.shape) ^ flips).astype(np.uint8))
        if case == "patches":
            gts[-1][:] = 0
        with tempfile.TemporaryDirectory() as tmp:
            out = helpers.run_reference_tta(tmp, case, jpegs=[first[f"jpeg{i}"] for i in range(n)], gts=gts)
        for i in range(n):
            assert np.array_equal(out[f"pred{i}"], first[f"pred{i}"])
            assert np.array_equal(out[f"gt{i}" if case in helpers.VESSEL_CASES else f"label{i}"], gts[i])
        np.savez_compressed(os.path.join(HERE, f"tta_{case}.npz"), **out)
        print(f"tta_{case}.npz written: auc", float(out["auc"]), "thresholds", out["thresholds"].tolist(),
              "mask means", [float(out[f"mask{i}"].mean()) for i in range(n)])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "tta":
        return tta_fixtures_only()
    if len(sys.argv) > 1 and sys.argv[1] == "pad":     # only the fixtures added in round 2
        return pad_fixture_only()
    if len(sys.argv) > 1 and sys.argv[1] == "ensemble":
        return ensemble_fixture_only()
    ref = ref_loader.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # 1. network outputs of the reference modules on product-initialised weights
    nets_out, layout = {}, {}
    for key, name, cfg, size, batch in helpers.NET_GOLDEN_CASES:
        product = helpers.build_product_model(name, cfg)
        module = reference_module(ref, name, cfg).eval()
        module.load_state_dict(product.state_dict(), strict=True)
        x = helpers.golden_input(batch, size)
        with torch.no_grad():
            y = module(x)
        nets_out[key] = y.numpy().astype(np.float32)
        layout[key] = {k: list(v.shape) for k, v in module.state_dict().items()}
        print(key, tuple(y.shape), float(y.mean()), float(y.std()))
    np.savez_compressed(os.path.join(HERE, "net_logits.npz"), **nets_out)
    json.dump(layout, open(os.path.join(HERE, "state_dict_layout.json"), "w"))

    # 2. scoring: the reference's aucpr.py on seeded synthetic probability maps
    scoring = {}
    cfg = {"out_dir": "/tmp/eds_golden_out", "dataset_name": "IDRiD", "lesion_type": "EX"}
    for seed in (0, 1, 2):
        items = helpers.synth_scoring_case(seed)
        scoring[str(seed)] = {
            "get_auc": float(ref.aucpr.get_auc(items, cfg)),
            "get_aucroc": float(ref.aucpr.get_aucroc(items, cfg)),
            "plot_aucpr_curve": [float(t) for t in ref.aucpr.plot_aucpr_curve(items, "golden", cfg)],
            "plot_aucroc_curve": float(ref.aucpr.plot_aucroc_curve(items, "golden", cfg)),
        }
    json.dump(scoring, open(os.path.join(HERE, "scoring.json"), "w"), indent=1)

    # 3. tiling
    shapes = [((2848, 4288), 2048, 32), ((2848, 4288), 1024, 32), ((608, 608), 512, 32), ((1024, 1024), 1024, 32),
              ((584, 565), 1024, 32), ((960, 999), 512, 32), ((2000, 3000), 256, 32), ((512, 512), 256, 0)]
    grids = [{"shape": list(s), "window": w, "min_overlap": o,
              "grid": ref.base_utils.make_grid(s, window=w, min_overlap=o).tolist()} for s, w, o in shapes]
    json.dump(grids, open(os.path.join(HERE, "make_grid.json"), "w"))

    # 4. preprocessing + registry names
    pre = {}
    sample = np.arange(0, 256, 5, dtype=np.uint8).reshape(-1, 1, 1).repeat(3, axis=2)
    for ds in ("IDRiD", "FGADR", "DDR", "DRIVE", "HRF", "CHASEDB1", None):
        fn, mean, std = ref_loader.get_preprocessing_fn(ds, False)
        pre[str(ds)] = {"mean": mean, "std": std, "out": fn(sample).astype(np.float32).reshape(-1).tolist()}
    json.dump({"preprocessing": pre, "registry": ref_loader.registry_names()},
              open(os.path.join(HERE, "registry_preprocessing.json"), "w"))
    # 5. stat_result.export_result (lesion + vessel twins) of the reference on a seeded mask set
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        lesion_cfg, vessel_cfg = helpers.make_stat_case(tmp, seed=0)
        ref.stat_result.export_result("EX/exp", lesion_cfg)
        ref.stat_result_vessel.export_result("vexp", vessel_cfg)
        stat = {"lesion": helpers.read_stat_csvs(os.path.join(tmp, "out", "IDRiD", "result_assessment", "EX", "exp")),
                "vessel": helpers.read_stat_csvs(os.path.join(tmp, "out", "DRIVE", "result_assessment", "vexp"))}
    json.dump(stat, open(os.path.join(HERE, "stat_result.json"), "w"), indent=1)
    # 6. vessel padding: the reference's own pad_img.pad, file to file, on tiny seeded images / labels
    pad_fixture_only()
    # 7. the reference's ensemble.py end to end
    ensemble_fixture_only()
    # 8. the reference's tta.py drivers end to end
    tta_fixtures_only()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
