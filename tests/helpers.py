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
