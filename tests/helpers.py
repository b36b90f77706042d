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
