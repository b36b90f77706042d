"""Two ranks on two GPUs (skipped on a one-GPU box): ``tta_patches`` launched as ``torchrun --nproc-per-node 2``
with the (image, tile) partition of SURVEY.md 8e against the same call in ONE process -- same files in, same
AUC-PR out (the integer histograms are all-reduced, so only the fp32-atomic jitter of the SE / SCSE channel means
separates the two runs), same masks on disk, every mask written exactly once."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np, torch
from eyediseasesegmentation_b200 import tta as eds_tta, _driver as drv
seen = {{}}
real = eds_tta.get_auc
eds_tta.get_auc = lambda gen, config: seen.setdefault('pr', real(gen, config))
cfg = json.load(open({cfg!r}))
config = dict(cfg['config'])
for k in ('test_img_path', 'test_mask_path'):
    config[k] = Path(config[k])
config['out_dir'] = {out!r}
eds_tta.tta_patches(cfg['logdir'], config, cfg['args'])
rank = int(os.environ.get('RANK', '0'))
json.dump(seen, open({out!r} + f'/auc_rank{{rank}}.json', 'w'))
import torch.distributed as dist
if dist.is_initialized():
    dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tta_patches_two_ranks_equal_one_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    name, cfg = "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)
    model = helpers.build_product_model(name, cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["segmentation_head.0.weight"] *= 6.0
    logdir = tmp_path / "models" / "IDRiD" / "EX" / "part"
    (logdir / "checkpoints").mkdir(parents=True)
    torch.save({"model_state_dict": sd}, logdir / "checkpoints" / "best.pth")
    img_dir, mask_root = tmp_path / "data" / "images", tmp_path / "data" / "masks"
    mask_dir = mask_root / "3. Hard Exudates"
    img_dir.mkdir(parents=True)
    mask_dir.mkdir(parents=True)
    rng = np.random.default_rng(1)
    shapes = [(300, 420), (280, 302), (330, 290)]          # three images, two ranks: a ragged last group
    for i, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        Image.fromarray(img).save(img_dir / f"IDRiD_{i:02d}.jpg", quality=95)
        gt = (np.kron(rng.random((h // 10 + 1, w // 10 + 1)) < 0.1, np.ones((10, 10)))[:h, :w] * 255).astype(np.uint8)
        Image.fromarray(gt, "L").save(mask_dir / f"IDRiD_{i:02d}_EX.tif")
    config = {"dataset_name": "IDRiD", "lesion_type": "EX", "gray": False, "scale_size": 128, "val_batch_size": 2,
              "model_name": name, "model_params": dict(cfg), "test_img_path": str(img_dir),
              "test_mask_path": str(mask_root), "data_type": "tile"}
    cfg_path = tmp_path / "cfg.json"
    json.dump({"config": config, "logdir": str(logdir),
               "args": {"best": "true", "tta": "d4", "createprob": "false", "optim_thres": 0}}, open(cfg_path, "w"))
    outs = {}
    for tag, launcher in (("one", [sys.executable]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200)])):
        out = tmp_path / f"out_{tag}"
        out.mkdir()
        script = tmp_path / f"worker_{tag}.py"
        script.write_text(_WORKER.format(root=ROOT, cfg=str(cfg_path), out=str(out)))
        env = dict(os.environ, EDS_PRECISION="fp32")
        res = subprocess.run(launcher + [str(script)], capture_output=True, text=True, env=env, timeout=900)
        assert res.returncode == 0, res.stderr[-3000:]
        outs[tag] = out
    one = json.load(open(outs["one"] / "auc_rank0.json"))["pr"]
    for r in range(2):
        assert abs(json.load(open(outs["two"] / f"auc_rank{r}.json"))["pr"] - one) < 1e-5      # every rank: global score
    d1, d2 = outs["one"] / "IDRiD" / "tta" / "EX" / "part", outs["two"] / "IDRiD" / "tta" / "EX" / "part"
    assert sorted(p.name for p in d1.iterdir()) == sorted(p.name for p in d2.iterdir()) == [f"IDRiD_{i:02d}.jpg" for i in range(3)]
    for p in d1.iterdir():
        a, b = np.asarray(Image.open(p).convert("L")) > 127, np.asarray(Image.open(d2 / p.name).convert("L")) > 127
        assert a.shape == b.shape and np.mean(a != b) < 1e-3, p.name
