"""TEST INFRASTRUCTURE: plain-PyTorch fp32 restatement of the three networks on the hot path.

Each function evaluates one reference module directly from a ``state_dict`` with the
reference's key layout, in eval mode, NCHW fp32, with no nn.Module
state.  tests/test_oracle.py pins these functions against the reference's own modules
(loaded by ``oracle/ref_loader.py``) bit-for-bit on CPU, and tests/golden/ holds outputs of
the reference itself for the GPU box, where /root/reference does not exist.

Citations are ``file:line`` under /root/reference/src/main/archs unless stated; "3P" marks
third-party code restated from its published architecture.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------ primitives
def bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm{1,2}d in eval mode (eps 1e-5)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def scse(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """3P smp SCSEModule: x * cSE(x) + x * sSE(x)."""
    z = F.adaptive_avg_pool2d(x, 1)
    z = F.relu(conv(sd, p + ".cSE.1", z))
    c = torch.sigmoid(conv(sd, p + ".cSE.3", z))
    s = torch.sigmoid(conv(sd, p + ".sSE.0", x))
    return x * c + x * s


def attention(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """3P smp Attention(name): SCSE when its parameters exist, Identity otherwise."""
    return scse(sd, p + ".attention", x) if (p + ".attention.sSE.0.weight") in sd else x


# ---------------------------------------------------------------------- SENet
def se_bottleneck(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """3P SEResNetBottleneck (stride on conv1) + SEModule, then residual add and ReLU."""
    out = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=stride)))
    out = F.relu(bn(sd, p + ".bn2", conv(sd, p + ".conv2", out, padding=1)))
    out = bn(sd, p + ".bn3", conv(sd, p + ".conv3", out))
    residual = x
    if (p + ".downsample.0.weight") in sd:
        residual = bn(sd, p + ".downsample.1", conv(sd, p + ".downsample.0", x, stride=stride))
    z = F.adaptive_avg_pool2d(out, 1)
    z = torch.sigmoid(conv(sd, p + ".se_module.fc2", F.relu(conv(sd, p + ".se_module.fc1", z))))
    return F.relu(out * z + residual)


def _n_blocks(sd: SD, p: str) -> int:
    n = 0
    while f"{p}.{n}.conv1.weight" in sd:
        n += 1
    return n


def senet_layer(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    for i in range(_n_blocks(sd, p)):
        x = se_bottleneck(sd, f"{p}.{i}", x, stride if i == 0 else 1)
    return x


def senet_stem(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """layer0 without its pool: conv1 7x7 s2 p3, bn1, relu1."""
    return F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=2, padding=3)))


# ------------------------------------------------------------- axial attention
def _relative(sd: SD, p: str, dim: int, d_kq: int, d_v: int):
    """axial_attention_v2.py:30-46: r[c, x, y] = relative[c, x - y + dim - 1]."""
    rel = sd[p + ".relative"]
    idx = (torch.arange(dim).view(dim, 1) - torch.arange(dim).view(1, dim) + dim - 1).to(rel.device)
    emb = rel[:, idx.reshape(-1)].reshape(rel.shape[0], dim, dim)
    return emb[:d_kq], emb[d_kq:2 * d_kq], emb[2 * d_kq:2 * d_kq + d_v]


def _conv1d_bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """_conv1d1x1 (axial_attention_v2.py:49-52): Conv1d(k=1, no bias) + BatchNorm1d."""
    return bn(sd, p + ".1", F.conv1d(x, sd[p + ".0.weight"]))


def _attend(sd: SD, p: str, q, k, v, dim: int, heads: int, d_kq: int, d_v: int) -> torch.Tensor:
    """Shared tail of AxialAttention.forward (:178-213) and CrossAxialAttention.forward (:100-135).
    q, k: [b, heads, d_kq, dim]; v: [b, heads, d_v, dim] -> [b, heads*d_v, dim]."""
    r_q, r_k, r_v = _relative(sd, p + ".RelativePosEncQKV", dim, d_kq, d_v)
    qr = torch.einsum("bhid,idj->bhdj", q, r_q)
    kr = torch.einsum("bhid,idj->bhdj", k, r_k)
    dots = torch.einsum("bhid,bhij->bhdj", q, k)
    b = q.shape[0]
    # 'n b h d j -> b (h n) d j' : channel = h * 3 + n, n in (qr, kr, dots)
    stacked = torch.stack([qr, kr, dots], dim=2).reshape(b, heads * 3, dim, dim)
    normed = bn(sd, p + ".attention_norm", stacked).reshape(b, heads, 3, dim, dim)
    attn = torch.softmax(normed.sum(dim=2), dim=-1)
    out = torch.einsum("bhdj,bhij->bhid", attn, v)
    kv = torch.einsum("bhdj,idj->bhid", attn, r_v)
    # 'n b h i d -> b (n h i) d', n in (kv, out)
    both = torch.stack([kv, out], dim=1).reshape(b, 2 * heads * d_v, dim)
    both = bn(sd, p + ".out_norm", both).reshape(b, 2, heads * d_v, dim)
    return both.sum(dim=1)


def axial_attention(sd: SD, p: str, x: torch.Tensor, dim: int, heads: int = 8, d_kq: int = 8) -> torch.Tensor:
    """AxialAttention.forward (axial_attention_v2.py:167-213); x: [b, C, dim]."""
    b, C, _ = x.shape
    d_v = C // heads
    qkv = _conv1d_bn(sd, p + ".to_qvk", x)
    # 'b (q h) d -> b h q d': channel = q * heads + h
    qkv = qkv.reshape(b, 2 * d_kq + d_v, heads, dim).permute(0, 2, 1, 3)
    q, k, v = qkv[:, :, :d_kq], qkv[:, :, d_kq:2 * d_kq], qkv[:, :, 2 * d_kq:]
    return _attend(sd, p, q, k, v, dim, heads, d_kq, d_v)


def cross_axial_attention(sd: SD, p: str, x_in: torch.Tensor, skip: torch.Tensor, dim: int, heads: int = 4,
                          d_kq: int = 8) -> torch.Tensor:
    """CrossAxialAttention.forward (axial_attention_v2.py:87-135); x_in: [b, Cx, dim], skip: [b, Cs, dim]."""
    b = x_in.shape[0]
    d_v = skip.shape[1] // heads
    qk = _conv1d_bn(sd, p + ".to_kq", x_in).reshape(b, 2 * d_kq, heads, dim).permute(0, 2, 1, 3)
    v = _conv1d_bn(sd, p + ".to_v", skip).reshape(b, d_v, heads, dim).permute(0, 2, 1, 3)
    return _attend(sd, p, qk[:, :, :d_kq], qk[:, :, d_kq:], v, dim, heads, d_kq, d_v)


def _rows_as_seq(x):   # 'b c h w -> (b w) c h'
    b, c, h, w = x.shape
    return x.permute(0, 3, 1, 2).reshape(b * w, c, h)


def _cols_as_seq(x):   # 'b c h w -> (b h) c w'
    b, c, h, w = x.shape
    return x.permute(0, 2, 1, 3).reshape(b * h, c, w)


def axial_block(sd: SD, p: str, x_in: torch.Tensor, dim: int, heads: int = 8) -> torch.Tensor:
    """AxialAttentionBlock.forward (axial_attention_v2.py:261-281)."""
    b = x_in.shape[0]
    x = F.relu(bn(sd, p + ".in_conv1x1.1", conv(sd, p + ".in_conv1x1.0", x_in)))
    c = x.shape[1]
    x = axial_attention(sd, p + ".height_att", _rows_as_seq(x), dim, heads)          # [(b w), c, h]
    x = x.reshape(b, dim, c, dim).permute(0, 3, 2, 1).reshape(b * dim, c, dim)       # '(b w) c h -> (b h) c w'
    x = axial_attention(sd, p + ".width_att", x, dim, heads)                         # [(b h), c, w]
    x = x.reshape(b, dim, c, dim).permute(0, 2, 1, 3)                                # '(b h) c w -> b c h w'
    if (p + ".shortcut.0.weight") in sd:                                             # down_sample=True
        x_in = bn(sd, p + ".shortcut.1", conv(sd, p + ".shortcut.0", x_in, stride=2, padding=1))
        x = bn(sd, p + ".att_down.1", F.avg_pool2d(x, 2))
    x = F.relu(x)
    out = bn(sd, p + ".out_conv1x1.1", conv(sd, p + ".out_conv1x1.0", x))
    return F.relu(out + x_in)


# ------------------------------------------------------- proposed net (UNet++*)
def star_encoder(sd: SD, x: torch.Tensor, base_dim: int) -> List[torch.Tensor]:
    """BoTSER50.forward (unetplusplusstar.py:341-352) with use_axial=True."""
    p = "encoder"
    f1 = senet_stem(sd, p + ".layer0", x)
    y = F.max_pool2d(f1, 3, stride=2, ceil_mode=True)          # applied after f1 is recorded (:347-348)
    f2 = senet_layer(sd, p + ".layer1", y, 1)
    f3 = senet_layer(sd, p + ".layer2", f2, 2)
    f4 = senet_layer(sd, p + ".layer3", f3, 2)
    y = axial_block(sd, p + ".layer4.0", f4, base_dim * 2)
    y = axial_block(sd, p + ".layer4.1", y, base_dim)
    f5 = axial_block(sd, p + ".layer4.2", y, base_dim)         # same module object as layer4.1 (:323-328)
    return [x, f1, f2, f3, f4, f5]


def _conv_bn_relu(sd: SD, p: str, x: torch.Tensor, bn_idx: int) -> torch.Tensor:
    """Conv2dReLU: 3x3 pad 1 conv -> (DropBlock: identity in eval) -> BN -> ReLU.
    bn_idx = 2 for unetplusplusstar.py:22-63, 1 for 3P smp Conv2dReLU."""
    return F.relu(bn(sd, f"{p}.{bn_idx}", conv(sd, p + ".0", x, padding=1)))


def star_decoder_block(sd: SD, p: str, x: torch.Tensor, skip: Optional[torch.Tensor], dim: int) -> torch.Tensor:
    """DecoderBlock.forward (unetplusplusstar.py:127-161)."""
    x_up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    use_catt = (p + ".down_sample.weight") in sd
    if skip is not None:
        if use_catt:
            b = x.shape[0]
            ori = F.conv2d(skip, sd[p + ".down_sample.weight"])
            s = F.max_pool2d(skip, 2)
            s = F.relu(bn(sd, p + ".init_conv.2", conv(sd, p + ".init_conv.1", s)))
            c = s.shape[1]
            s = cross_axial_attention(sd, p + ".h_catt", _rows_as_seq(x), _rows_as_seq(s), dim)
            s = s.reshape(b, dim, c, dim).permute(0, 3, 2, 1).reshape(b * dim, c, dim)
            s = cross_axial_attention(sd, p + ".w_catt", _cols_as_seq(x), s, dim)
            s = s.reshape(b, dim, c, dim).permute(0, 2, 1, 3)
            gate = F.interpolate(torch.sigmoid(s), scale_factor=2, mode="bilinear", align_corners=False)
            s = F.conv2d(ori * gate, sd[p + ".up_sample.weight"])
            x_up = torch.cat([x_up, s], dim=1)
        else:
            x_up = attention(sd, p + ".attention1", torch.cat([x_up, skip], dim=1))
    y = _conv_bn_relu(sd, p + ".conv1", x_up, 2)
    y = _conv_bn_relu(sd, p + ".conv2", y, 2)
    if not use_catt:
        y = attention(sd, p + ".attention2", y)
    return y


def _dense_decoder(features: List[torch.Tensor], block_fn) -> torch.Tensor:
    """UnetPlusPlusDecoder.forward (unetplusplusstar.py:239-263 == deep_supunetplusplus.py:116-139).
    block_fn(name, layer_idx, x, skip) evaluates decoder block `name`."""
    feats = features[1:][::-1]
    depth = len(feats) - 1
    dense = {}
    for layer_idx in range(depth):
        for depth_idx in range(depth - layer_idx):
            if layer_idx == 0:
                name = f"x_{depth_idx}_{depth_idx}"
                dense[name] = block_fn(name, depth_idx, feats[depth_idx], feats[depth_idx + 1])
            else:
                li = depth_idx + layer_idx
                cat = [dense[f"x_{i}_{li}"] for i in range(depth_idx + 1, li + 1)] + [feats[li + 1]]
                name = f"x_{depth_idx}_{li}"
                dense[name] = block_fn(name, li, dense[f"x_{depth_idx}_{li - 1}"], torch.cat(cat, dim=1))
    name = f"x_0_{depth}"
    dense[name] = block_fn(name, 0, dense[f"x_0_{depth - 1}"], None)
    return dense[name]


def unetplusplusstar_forward(sd: SD, x: torch.Tensor, base_dim: int, return_features: bool = False):
    """UnetPlusPlusStar.forward (unetplusplusstar.py:465-488), deep_supervision=False, clf_head=False."""
    feats = star_encoder(sd, x, base_dim)

    def block(name, level, xx, skip):
        return star_decoder_block(sd, f"decoder.blocks.{name}", xx, skip, base_dim * (2 ** level))

    y = _dense_decoder(feats, block)
    mask = conv(sd, "segmentation_head.0", y, padding=1)
    return (mask, feats) if return_features else mask


# ------------------------------------------------ baseline UNet++ and smp.Unet
def resnet_basic_block(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    out = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=stride, padding=1)))
    out = bn(sd, p + ".bn2", conv(sd, p + ".conv2", out, padding=1))
    idt = x
    if (p + ".downsample.0.weight") in sd:
        idt = bn(sd, p + ".downsample.1", conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(out + idt)


def smp_encoder(sd: SD, x: torch.Tensor) -> List[torch.Tensor]:
    """3P smp encoders: SENetEncoder (se_resnet50) or ResNetEncoder (resnet34), by key layout."""
    p = "encoder"
    if (p + ".layer0.conv1.weight") in sd:
        f1 = senet_stem(sd, p + ".layer0", x)
        y = F.max_pool2d(f1, 3, stride=2, ceil_mode=True)
        f2 = senet_layer(sd, p + ".layer1", y, 1)
        f3 = senet_layer(sd, p + ".layer2", f2, 2)
        f4 = senet_layer(sd, p + ".layer3", f3, 2)
        f5 = senet_layer(sd, p + ".layer4", f4, 2)
        return [x, f1, f2, f3, f4, f5]
    f1 = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=2, padding=3)))
    y = F.max_pool2d(f1, 3, stride=2, padding=1)
    feats = [x, f1]
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for i in range(_n_blocks(sd, f"{p}.layer{li}")):
            y = resnet_basic_block(sd, f"{p}.layer{li}.{i}", y, stride if i == 0 else 1)
        feats.append(y)
    return feats


def smp_decoder_block(sd: SD, p: str, x: torch.Tensor, skip: Optional[torch.Tensor]) -> torch.Tensor:
    """deep_supunetplusplus.py:48-56 / 3P smp unet DecoderBlock: nearest x2, cat, attention1, conv1, conv2, attention2."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    if skip is not None:
        x = attention(sd, p + ".attention1", torch.cat([x, skip], dim=1))
    x = _conv_bn_relu(sd, p + ".conv1", x, 1)
    x = _conv_bn_relu(sd, p + ".conv2", x, 1)
    return attention(sd, p + ".attention2", x)


def unetplusplus_forward(sd: SD, x: torch.Tensor, return_features: bool = False):
    """deep_supunetplusplus.UnetPlusPlus.forward (:258-273) with deep_supervision=False."""
    feats = smp_encoder(sd, x)
    y = _dense_decoder(feats, lambda name, level, xx, skip: smp_decoder_block(sd, f"decoder.blocks.{name}", xx, skip))
    mask = conv(sd, "segmentation_head.0", y, padding=1)
    return (mask, feats) if return_features else mask


def unet_forward(sd: SD, x: torch.Tensor, return_features: bool = False):
    """3P smp.Unet.forward: 5 decoder blocks, skips = reversed encoder features, no center block."""
    feats = smp_encoder(sd, x)
    rev = feats[1:][::-1]
    y = rev[0]
    for i in range(5):
        y = smp_decoder_block(sd, f"decoder.blocks.{i}", y, rev[i + 1] if i + 1 < len(rev) else None)
    mask = conv(sd, "segmentation_head.0", y, padding=1)
    return (mask, feats) if return_features else mask


def forward(model_name: str, sd: SD, x: torch.Tensor, params: Optional[dict] = None) -> torch.Tensor:
    """Dispatch by the reference's registry / smp name."""
    if model_name == "unetplusplusstar":
        return unetplusplusstar_forward(sd, x, int((params or {}).get("base_dim", 32)))
    if model_name == "unetplusplus_deepsup":
        return unetplusplus_forward(sd, x)
    if model_name == "Unet":
        return unet_forward(sd, x)
    raise KeyError(model_name)


# ------------------------------------------------------------------------ TTA
def tta_views(kind: str):
    """3P ttach 0.0.3 aliases: list of (augment, deaugment) closures in view order."""
    def hflip(t): return t.flip(3)
    def vflip(t): return t.flip(2)
    def rot(k): return lambda t: torch.rot90(t, k, (2, 3))
    ident = lambda t: t  # noqa: E731
    if kind == "d4":       # HorizontalFlip x Rotate90(0, 90, 180, 270)
        views = []
        for flip in (False, True):
            for k in range(4):
                aug = (lambda t, f=flip, kk=k: rot(kk)(hflip(t) if f else t))
                deaug = (lambda t, f=flip, kk=k: (hflip(rot((4 - kk) % 4)(t)) if f else rot((4 - kk) % 4)(t)))
                views.append((aug, deaug))
        return views
    if kind == "flip":     # HorizontalFlip x VerticalFlip
        views = []
        for hf in (False, True):
            for vf in (False, True):
                aug = (lambda t, a=hf, b=vf: (vflip if b else ident)((hflip if a else ident)(t)))
                deaug = (lambda t, a=hf, b=vf: (hflip if a else ident)((vflip if b else ident)(t)))
                views.append((aug, deaug))
        return views
    if kind == "hflip":
        return [(ident, ident), (hflip, hflip)]
    if kind == "none":
        return [(ident, ident)]
    raise KeyError(kind)


def tta_mean_logits(net, x: torch.Tensor, kind: str) -> torch.Tensor:
    """ttach.SegmentationTTAWrapper(model, <kind>_transform(), merge_mode='mean') (tta.py:92-99)."""
    views = tta_views(kind)
    total = None
    for aug, deaug in views:
        y = deaug(net(aug(x)))
        total = y if total is None else total + y
    return total / len(views)
