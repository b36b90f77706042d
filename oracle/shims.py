"""TEST INFRASTRUCTURE -- not part of the product path.

nn.Module restatements of the third-party pieces the reference's hot-path files import
but that are neither vendored under /root/reference nor installed here:

  pytorch-toolbelt 0.4.2  ``modules.backbone.senet.se_resnet50``   (Cadene SENet)
  segmentation-models-pytorch 0.1.3  ``base.modules`` (Attention / SCSEModule / Activation /
        Flatten / Conv2dReLU), ``base`` (SegmentationModel / SegmentationHead /
        ClassificationHead / initialization), ``encoders.get_encoder`` (resnet34,
        se_resnet50), ``Unet``
  timm 0.3.2  ``models.layers.DropBlock2d`` / ``DropPath`` (identity in eval mode)

They exist so that ``oracle/ref_loader.py`` can import the reference's own arch files by
path in the build container, and so that state_dict key layouts can be pinned.  They are
written from the published architectures of those packages; parity with the real packages
is unpinned (the packages cannot be installed offline) apart from the parameter counts
recorded in SURVEY.md 8c, which tests/test_oracle.py re-checks.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn
import torch.nn.functional as F


# ----------------------------------------------------------------------- SENet
class SEModule(nn.Module):
    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        s = self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))
        return x * s


class SEResNetBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, groups, reduction, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False, stride=stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.se_module = SEModule(planes * 4, reduction=reduction)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        return self.relu(self.se_module(out) + residual)


class SENet(nn.Module):
    def __init__(self, block, layers, groups, reduction, inplanes=64, downsample_kernel_size=1,
                 downsample_padding=0, num_classes=1000):
        super().__init__()
        self.inplanes = inplanes
        self.layer0 = nn.Sequential(OrderedDict([
            ("conv1", nn.Conv2d(3, inplanes, kernel_size=7, stride=2, padding=3, bias=False)),
            ("bn1", nn.BatchNorm2d(inplanes)),
            ("relu1", nn.ReLU(inplace=True)),
            ("pool", nn.MaxPool2d(3, stride=2, ceil_mode=True)),
        ]))
        kw = dict(groups=groups, reduction=reduction, downsample_kernel_size=downsample_kernel_size,
                  downsample_padding=downsample_padding)
        self.layer1 = self._make_layer(block, 64, layers[0], stride=1, **kw)
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, **kw)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, **kw)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, **kw)
        self.avg_pool = nn.AvgPool2d(7, stride=1)
        self.dropout = None
        self.last_linear = nn.Linear(512 * block.expansion, num_classes)

    def _make_layer(self, block, planes, blocks, groups, reduction, stride, downsample_kernel_size,
                    downsample_padding):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=downsample_kernel_size, stride=stride,
                          padding=downsample_padding, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, groups, reduction, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, groups, reduction))
        return nn.Sequential(*layers)


def se_resnet50(pretrained=None, num_classes=1000):
    return SENet(SEResNetBottleneck, [3, 4, 6, 3], groups=1, reduction=16, num_classes=num_classes)


# ------------------------------------------------------- smp.base.modules (md)
class Conv2dReLU(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, use_batchnorm=True):
        conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         bias=not use_batchnorm)
        bn = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        super().__init__(conv, bn, nn.ReLU(inplace=True))


class SCSEModule(nn.Module):
    def __init__(self, in_channels, reduction=16):
        super().__init__()
        self.cSE = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(in_channels, in_channels // reduction, 1),
            nn.ReLU(inplace=True),
            nn.Conv2d(in_channels // reduction, in_channels, 1),
            nn.Sigmoid())
        self.sSE = nn.Sequential(nn.Conv2d(in_channels, 1, 1), nn.Sigmoid())

    def forward(self, x):
        return x * self.cSE(x) + x * self.sSE(x)


class Attention(nn.Module):
    def __init__(self, name, **params):
        super().__init__()
        if name is None:
            self.attention = nn.Identity(**params)
        elif name == "scse":
            self.attention = SCSEModule(**params)
        else:
            raise ValueError("Attention {} is not implemented".format(name))

    def forward(self, x):
        return self.attention(x)


class Activation(nn.Module):
    def __init__(self, name, **params):
        super().__init__()
        if name is None or name == "identity":
            self.activation = nn.Identity(**params)
        elif name == "sigmoid":
            self.activation = nn.Sigmoid()
        elif callable(name):
            self.activation = name(**params)
        else:
            raise ValueError("activation {} not restated".format(name))

    def forward(self, x):
        return self.activation(x)


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.shape[0], -1)


class SegmentationHead(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size=3, activation=None, upsampling=1):
        conv2d = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2)
        up = nn.UpsamplingBilinear2d(scale_factor=upsampling) if upsampling > 1 else nn.Identity()
        super().__init__(conv2d, up, Activation(activation))


class ClassificationHead(nn.Sequential):
    def __init__(self, in_channels, classes, pooling="avg", dropout=0.2, activation=None):
        pool = nn.AdaptiveAvgPool2d(1) if pooling == "avg" else nn.AdaptiveMaxPool2d(1)
        drop = nn.Dropout(p=dropout, inplace=True) if dropout else nn.Identity()
        super().__init__(pool, Flatten(), drop, nn.Linear(in_channels, classes, bias=True), Activation(activation))


class SegmentationModel(nn.Module):
    def initialize(self):
        initialize_decoder(self.decoder)
        initialize_head(self.segmentation_head)
        if getattr(self, "classification_head", None) is not None:
            initialize_head(self.classification_head)

    def forward(self, x):
        features = self.encoder(x)
        masks = self.segmentation_head(self.decoder(*features))
        if getattr(self, "classification_head", None) is not None:
            return masks, self.classification_head(features[-1])
        return masks


def initialize_decoder(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


def initialize_head(module):
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


# ----------------------------------------------------------------- smp encoders
class _SENetEncoder(SENet):
    """smp SENetEncoder: stages identity | layer0[:-1] | pool+layer1 | layer2 | layer3 | layer4."""

    def __init__(self, depth=5):
        super().__init__(SEResNetBottleneck, [3, 4, 6, 3], groups=1, reduction=16)
        self.out_channels = (3, 64, 256, 512, 1024, 2048)
        self._depth = depth
        del self.last_linear
        del self.avg_pool

    def forward(self, x):
        stages = [nn.Identity(), self.layer0[:-1], nn.Sequential(self.layer0[-1], self.layer1), self.layer2,
                  self.layer3, self.layer4]
        feats = []
        for i in range(self._depth + 1):
            x = stages[i](x)
            feats.append(x)
        return feats


class _BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + identity)


class _ResNet34Encoder(nn.Module):
    """smp ResNetEncoder(resnet34): torchvision ResNet minus fc/avgpool; stages identity |
    conv1,bn1,relu | maxpool,layer1 | layer2 | layer3 | layer4."""

    def __init__(self, depth=5):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.inplanes = 64
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        self.out_channels = (3, 64, 64, 128, 256, 512)
        self._depth = depth
        for m in self.modules():  # torchvision's default init
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [_BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        layers += [_BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        stages = [nn.Identity(), nn.Sequential(self.conv1, self.bn1, self.relu),
                  nn.Sequential(self.maxpool, self.layer1), self.layer2, self.layer3, self.layer4]
        feats = []
        for i in range(self._depth + 1):
            x = stages[i](x)
            feats.append(x)
        return feats


def get_encoder(name, in_channels=3, depth=5, weights=None):
    if in_channels != 3 or weights is not None:
        raise ValueError("oracle shim: only 3-channel, randomly initialised encoders are restated")
    if name == "se_resnet50":
        return _SENetEncoder(depth)
    if name == "resnet34":
        return _ResNet34Encoder(depth)
    raise KeyError("oracle shim: encoder {} is not restated".format(name))


# --------------------------------------------------------------------- smp.Unet
class _UnetDecoderBlock(nn.Module):
    def __init__(self, in_channels, skip_channels, out_channels, use_batchnorm=True, attention_type=None):
        super().__init__()
        self.conv1 = Conv2dReLU(in_channels + skip_channels, out_channels, 3, padding=1, use_batchnorm=use_batchnorm)
        self.attention1 = Attention(attention_type, in_channels=in_channels + skip_channels)
        self.conv2 = Conv2dReLU(out_channels, out_channels, 3, padding=1, use_batchnorm=use_batchnorm)
        self.attention2 = Attention(attention_type, in_channels=out_channels)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = self.attention1(torch.cat([x, skip], dim=1))
        return self.attention2(self.conv2(self.conv1(x)))


class _UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels, use_batchnorm=True, attention_type=None):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = list(enc[1:]) + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList([
            _UnetDecoderBlock(i, s, o, use_batchnorm=use_batchnorm, attention_type=attention_type)
            for i, s, o in zip(in_ch, skip_ch, decoder_channels)])

    def forward(self, *features):
        features = features[1:][::-1]
        x = self.center(features[0])
        skips = features[1:]
        for i, block in enumerate(self.blocks):
            x = block(x, skips[i] if i < len(skips) else None)
        return x


class Unet(SegmentationModel):
    def __init__(self, encoder_name="resnet34", encoder_depth=5, encoder_weights="imagenet",
                 decoder_use_batchnorm=True, decoder_channels=(256, 128, 64, 32, 16), decoder_attention_type=None,
                 in_channels=3, classes=1, activation=None, aux_params=None):
        super().__init__()
        self.encoder = get_encoder(encoder_name, in_channels=in_channels, depth=encoder_depth, weights=encoder_weights)
        self.decoder = _UnetDecoder(self.encoder.out_channels, decoder_channels, decoder_use_batchnorm,
                                    decoder_attention_type)
        self.segmentation_head = SegmentationHead(decoder_channels[-1], classes, activation=activation, kernel_size=3)
        self.classification_head = None
        self.name = "u-{}".format(encoder_name)
        self.initialize()


# ------------------------------------------------------------------------ timm
class DropBlock2d(nn.Module):
    """Identity outside training (the hot path always runs .eval())."""

    def __init__(self, drop_prob=0.1, block_size=7, **kwargs):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob:
            raise RuntimeError("oracle shim: DropBlock2d is only restated for eval mode")
        return x


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob:
            raise RuntimeError("oracle shim: DropPath is only restated for eval mode")
        return x
