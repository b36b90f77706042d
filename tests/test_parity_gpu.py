"""Network-level parity: the CUDA path (through the registry / model interface) against the
oracle (oracle/nets.py, pinned to the reference in tests/test_oracle.py) on the same
random-init weights and seeded inputs.

Tolerances (BASELINE.json north_star): probability maps within 1e-4 max-abs in fp32 mode and
2e-2 in bf16 mode.  Because random-init logits have a small spread, the bf16 runs are also
held to a relative-L2 bound on the logits and on every encoder feature map, which is the
check that actually catches structural mistakes.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from eyediseasesegmentation_b200 import archs, ttach_compat as tta  # noqa: E402
from oracle import nets  # noqa: E402

from helpers import randomize_bn, rel_l2, oracle_forward, star_cfg  # noqa: E402


def setup_module(module):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _input(b, s, seed=7):
    return torch.randn(b, 3, s, s, generator=torch.Generator().manual_seed(seed))


CASES = [
    ("unetplusplusstar", star_cfg(8), 256, 2),
    ("unetplusplusstar", star_cfg(16), 512, 1),
    ("unetplusplusstar", star_cfg(19), 608, 1),   # BASELINE config 5: DRIVE 584x565 padded to 608^2 (odd base_dim)
    ("unetplusplus_deepsup", dict(encoder_name="se_resnet50", encoder_weights=None, classes=1,
                                  decoder_attention_type="scse", deep_supervision=True), 256, 2),
    ("unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 256, 2),
    ("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1), 512, 1),  # BASELINE config 1
]


def _build(name, cfg, seed=1999):
    torch.manual_seed(seed)
    model = archs.Unet(**dict(cfg)) if name == "Unet" else archs.get_model(name, dict(cfg), training=False)
    randomize_bn(model, seed + 1)
    return model.eval()


@pytest.mark.parametrize("name,cfg,size,batch", CASES)
def test_fp32_mode_matches_oracle(name, cfg, size, batch):
    model = _build(name, cfg)
    x = _input(batch, size)
    ref, ref_feats = oracle_forward(name, cfg, model.state_dict(), x, features=True)
    model = model.to("cuda")
    model.precision = "fp32"
    eng = model.engine()
    eng.keep_features = True
    out = model(x.cuda()).cpu()
    assert out.shape == ref.shape
    for i in range(1, 6):
        f = eng.features[f"f{i}"].float().permute(0, 3, 1, 2).cpu()
        assert rel_l2(f, ref_feats[i]) < 1e-4, f"encoder feature f{i}"
    assert rel_l2(out, ref) < 1e-4
    assert (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item() < 1e-4  # north_star fp32 tolerance


@pytest.mark.parametrize("name,cfg,size,batch", CASES)
def test_bf16_mode_matches_oracle(name, cfg, size, batch):
    model = _build(name, cfg)
    x = _input(batch, size)
    ref, ref_feats = oracle_forward(name, cfg, model.state_dict(), x, features=True)
    model = model.to("cuda")
    model.precision = "bf16"
    eng = model.engine()
    eng.keep_features = True
    out = model(x.cuda()).cpu()
    for i in range(1, 6):
        f = eng.features[f"f{i}"].float().permute(0, 3, 1, 2).cpu()
        assert rel_l2(f, ref_feats[i]) < 3e-2, f"encoder feature f{i}: {rel_l2(f, ref_feats[i])}"
    assert rel_l2(out - out.mean(), ref - ref.mean()) < 8e-2, rel_l2(out - out.mean(), ref - ref.mean())
    assert (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item() < 2e-2  # north_star bf16 tolerance


@pytest.mark.parametrize("alias,kind", [("d4_transform", "d4"), ("flip_transform", "flip")])
def test_fused_tta_matches_ttach_semantics(alias, kind):
    name, cfg = "unetplusplusstar", star_cfg(8)
    model = _build(name, cfg)
    x = _input(2, 256, seed=11)
    sd = model.state_dict()
    ref = nets.tta_mean_logits(lambda t: nets.forward(name, sd, t, cfg), x, kind)
    model = model.to("cuda")
    model.precision = "fp32"
    wrapped = tta.SegmentationTTAWrapper(model, getattr(tta.aliases, alias)(), merge_mode="mean")
    out = wrapped(x.cuda()).cpu()
    assert out.shape == ref.shape
    assert (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item() < 1e-4
    # the unfused route (one forward per materialised view) must agree with the fused one
    plain = tta.Merger("mean", 8 if kind == "d4" else 4)
    for t in getattr(tta.aliases, alias)():
        plain.append(t.deaugment_mask(model(t.augment_image(x.cuda()).contiguous())))
    assert (plain.result.cpu() - out).abs().max().item() < 1e-4
    model.precision = "bf16"
    out16 = wrapped(x.cuda()).cpu()
    assert (torch.sigmoid(out16) - torch.sigmoid(ref)).abs().max().item() < 2e-2


def test_reference_checkpoint_roundtrip(tmp_path):
    """tta.py:86-87: torch.load(...)['model_state_dict'] -> load_state_dict -> identical outputs."""
    name, cfg = "unetplusplus_deepsup", dict(encoder_name="resnet34", encoder_weights=None, classes=1)
    a = _build(name, cfg, seed=3).to("cuda")
    torch.save({"model_state_dict": a.state_dict()}, tmp_path / "best.pth")
    b = _build(name, cfg, seed=4).to("cuda")
    x = _input(1, 256).cuda()
    assert not torch.equal(a(x), b(x))
    b.load_state_dict(torch.load(tmp_path / "best.pth")["model_state_dict"])
    assert torch.equal(a(x), b(x))


def test_no_cpu_fallback():
    model = _build("Unet", dict(encoder_name="resnet34", encoder_weights=None, classes=1))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(_input(1, 64))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_cuda_graph_replay_equals_eager(precision, monkeypatch):
    """forward_tta captures a CUDA graph on the second call of a shape; replays (with NEW inputs copied
    into the static buffer) must reproduce the eager launches (up to the run-to-run jitter of the fp32
    atomic accumulation of the SE / SCSE channel means), and two models sharing the graph memory pool
    must not disturb each other."""
    monkeypatch.setenv("EDS_CUDA_GRAPHS", "1")
    tfm = tta.aliases.d4_transform()
    models = []
    for seed in (1999, 2001):
        m = _build("unetplusplusstar", star_cfg(8), seed=seed).to("cuda")
        m.precision = precision
        models.append(m)
    xs = [_input(2, 256, seed=s).cuda() for s in (7, 8, 9)]
    monkeypatch.setenv("EDS_CUDA_GRAPHS", "0")
    want = [[m.forward_tta(x, tfm, True).clone() for x in xs] for m in models]
    monkeypatch.setenv("EDS_CUDA_GRAPHS", "1")
    for rep in range(2):                      # call 1 eager, call 2 captures + replays, then pure replays
        for mi, m in enumerate(models):
            for xi, x in enumerate(xs):
                got = m.forward_tta(x, tfm, True)
                err = (got - want[mi][xi]).abs().max().item()
                assert err < (5e-3 if precision == "bf16" else 1e-5), f"model {mi} input {xi} pass {rep}: {err}"
                # a wrong static-input copy or a clobbered pool would give another input's / model's answer
                other = (got - want[mi][(xi + 1) % 3]).abs().max().item()
                assert other > 10 * err + 1e-3
    assert all(any(isinstance(v, tuple) and v[3] > 100 for v in m._graphs.values()) for m in models), "no graph was captured"


def test_full_size_tile_matches_oracle():
    """BASELINE config 3 at its real size: the proposed network (base_dim 32) on one 1024 x 1024 tile with
    8-view D4 TTA -- every layer shape of the benchmark (K = 9216 reductions, 512^2 x 448-channel halo
    layers, attention over 64-long sequences) against the fp32 oracle: 1e-4 in fp32 mode, 2e-2 in bf16."""
    name, cfg = "unetplusplusstar", star_cfg(32)
    model = _build(name, cfg)
    sd = model.state_dict()
    x = _input(1, 1024, seed=5)
    with torch.no_grad():
        want = torch.sigmoid(nets.tta_mean_logits(lambda t: nets.forward(name, sd, t, cfg), x, "d4"))[0, 0]
    model = model.to("cuda")
    tfm = tta.aliases.d4_transform()
    for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        model.precision = precision
        got = model.forward_tta(x.cuda(), tfm, apply_sigmoid=True)[0, 0].cpu()
        err = (got - want).abs().max().item()
        assert err < tol, f"{precision}: max |dprob| {err}"
        assert rel_l2(got - got.mean(), want - want.mean()) < (1e-4 if precision == "fp32" else 8e-2)
