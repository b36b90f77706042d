"""Host-side mirror of the slice of ``ttach`` (pinned 0.0.3 by the reference's Pipfile) that
``src/main/tta.py:92-99,173-180`` and ``tta_vessel.py:90-97,161-168`` use:
``ttach.aliases.<name>_transform()`` and ``ttach.SegmentationTTAWrapper(model, transforms,
merge_mode="mean")``.

ttach is a third-party dependency that is not vendored in the reference tree, so this is a
restatement of its published behaviour (flip = ``x.flip(dim)``, rotate = ``torch.rot90(x, k,
(2, 3))``, views = ``itertools.product`` of the parameter lists in declaration order,
de-augmentation in reverse order with inverse parameters, Merger('mean') = running sum / n).

When the wrapped model is one of this package's B200 networks and every transform is a
flip / rot90, the wrapper does not materialise any augmented tensor: the view index maps go
to the stem kernel (``eds_stem_conv7x7s2``) and to the merge kernel (``eds_tta_merge``).
"""
from __future__ import annotations

import itertools
from functools import partial
from typing import List, Sequence, Tuple

import torch
from torch import nn


# ------------------------------------------------------------------ transforms
class BaseTransform:
    identity_param = None

    def __init__(self, name: str, params):
        self.params = params
        self.pname = name

    def apply_aug_image(self, image, *args, **params):
        raise NotImplementedError

    def apply_deaug_mask(self, mask, *args, **params):
        raise NotImplementedError


class HorizontalFlip(BaseTransform):
    identity_param = False

    def __init__(self):
        super().__init__("apply", [False, True])

    def apply_aug_image(self, image, apply=False, **kwargs):
        return image.flip(3) if apply else image

    def apply_deaug_mask(self, mask, apply=False, **kwargs):
        return mask.flip(3) if apply else mask


class VerticalFlip(BaseTransform):
    identity_param = False

    def __init__(self):
        super().__init__("apply", [False, True])

    def apply_aug_image(self, image, apply=False, **kwargs):
        return image.flip(2) if apply else image

    def apply_deaug_mask(self, mask, apply=False, **kwargs):
        return mask.flip(2) if apply else mask


class Rotate90(BaseTransform):
    identity_param = 0

    def __init__(self, angles: List[int]):
        if self.identity_param not in angles:
            angles = [self.identity_param] + list(angles)
        super().__init__("angle", list(angles))

    def apply_aug_image(self, image, angle=0, **kwargs):
        k = angle // 90 if angle >= 0 else (angle + 360) // 90
        return torch.rot90(image, k, (2, 3))

    def apply_deaug_mask(self, mask, angle=0, **kwargs):
        return self.apply_aug_image(mask, -angle)


class Scale(BaseTransform):
    """Multiscale views (``multiscale_transform``; tta.py:93-96 passes scales=[1, 2, 4])."""
    identity_param = 1

    def __init__(self, scales, interpolation: str = "nearest", align_corners=None):
        if self.identity_param not in scales:
            scales = [self.identity_param] + list(scales)
        self.interpolation = interpolation
        self.align_corners = align_corners
        super().__init__("scale", list(scales))

    def _resize(self, x, scale):
        return nn.functional.interpolate(x, scale_factor=scale, mode=self.interpolation,
                                         align_corners=self.align_corners)

    def apply_aug_image(self, image, scale=1, **kwargs):
        return image if scale == self.identity_param else self._resize(image, scale)

    def apply_deaug_mask(self, mask, scale=1, **kwargs):
        return mask if scale == self.identity_param else self._resize(mask, 1 / scale)


class Chain:
    def __init__(self, functions):
        self.functions = functions or []

    def __call__(self, x):
        for f in self.functions:
            x = f(x)
        return x


class Transformer:
    def __init__(self, image_pipeline: Chain, mask_pipeline: Chain, spec):
        self.image_pipeline = image_pipeline
        self.mask_pipeline = mask_pipeline
        self.spec = spec  # ((transform, param), ...) in augmentation order

    def augment_image(self, image):
        return self.image_pipeline(image)

    def deaugment_mask(self, mask):
        return self.mask_pipeline(mask)


class Compose:
    def __init__(self, transforms: Sequence[BaseTransform]):
        self.aug_transforms = list(transforms)
        self.aug_transform_parameters = list(itertools.product(*[t.params for t in self.aug_transforms]))
        self.deaug_transforms = self.aug_transforms[::-1]
        self.deaug_transform_parameters = [p[::-1] for p in self.aug_transform_parameters]

    def __iter__(self):
        for aug_params, deaug_params in zip(self.aug_transform_parameters, self.deaug_transform_parameters):
            image_chain = Chain([partial(t.apply_aug_image, **{t.pname: p})
                                 for t, p in zip(self.aug_transforms, aug_params)])
            mask_chain = Chain([partial(t.apply_deaug_mask, **{t.pname: p})
                                for t, p in zip(self.deaug_transforms, deaug_params)])
            yield Transformer(image_chain, mask_chain, tuple(zip(self.aug_transforms, aug_params)))

    def __len__(self):
        return len(self.aug_transform_parameters)


class Merger:
    def __init__(self, type: str = "mean", n: int = 1):
        if type not in ("mean", "gmean", "sum", "max", "min", "tsharpen"):
            raise ValueError(f"Not correct merge type `{type}`.")
        self.output = None
        self.type = type
        self.n = n

    def append(self, x):
        if self.type == "tsharpen":
            x = x ** 0.5
        if self.output is None:
            self.output = x
        elif self.type in ("mean", "sum", "tsharpen"):
            self.output = self.output + x
        elif self.type == "gmean":
            self.output = self.output * x
        elif self.type == "max":
            self.output = torch.max(self.output, x)
        elif self.type == "min":
            self.output = torch.min(self.output, x)

    @property
    def result(self):
        if self.type in ("sum", "max", "min"):
            return self.output
        if self.type in ("mean", "tsharpen"):
            return self.output / self.n
        return self.output ** (1 / self.n)


class _Aliases:
    @staticmethod
    def flip_transform():
        return Compose([HorizontalFlip(), VerticalFlip()])

    @staticmethod
    def hflip_transform():
        return Compose([HorizontalFlip()])

    @staticmethod
    def vlip_transform():  # sic: ttach spells it this way
        return Compose([VerticalFlip()])

    @staticmethod
    def d4_transform():
        return Compose([HorizontalFlip(), Rotate90(angles=[0, 90, 180, 270])])

    @staticmethod
    def multiscale_transform(scales, interpolation="nearest"):
        return Compose([Scale(scales, interpolation=interpolation)])


aliases = _Aliases()


# ---------------------------------------------------------- view index maps
def _fit_map(coords: torch.Tensor) -> Tuple[int, ...]:
    """coords [2,H,W] holding (row, col) of the source element -> (a,b,c,d,e,f) with
    row = a*i + b*j + c, col = d*i + e*j + f, verified on the whole grid."""
    H, W = coords.shape[1:]
    r, c = coords[0].long(), coords[1].long()
    i1, j1 = min(1, H - 1), min(1, W - 1)
    m = (int(r[i1, 0] - r[0, 0]), int(r[0, j1] - r[0, 0]), int(r[0, 0]),
         int(c[i1, 0] - c[0, 0]), int(c[0, j1] - c[0, 0]), int(c[0, 0]))
    ii = torch.arange(H).view(H, 1)
    jj = torch.arange(W).view(1, W)
    if not (torch.equal(m[0] * ii + m[1] * jj + m[2], r) and torch.equal(m[3] * ii + m[4] * jj + m[5], c)):
        raise ValueError("transform is not an axis-aligned flip / rot90")
    return m


def view_maps(transforms: Compose, H: int, W: int):
    """Index maps of every view, derived by pushing a coordinate grid through the transform
    chain itself (so they cannot drift from the semantics above).

    returns (aug, deaug): aug[v] maps augmented-image pixels to source pixels,
    deaug[v] maps merged-output pixels to pixels of view v's network output.
    """
    # cached on the Compose object: pushing two 1024^2 coordinate grids through 8 views costs tens of
    # milliseconds of host time (hundreds with OMP_NUM_THREADS=1 under torchrun) and never changes
    cache = transforms.__dict__.setdefault("_eds_view_maps", {})
    if (H, W) in cache:
        return cache[(H, W)]
    grid = torch.stack(torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")).float()[None]
    aug, deaug = [], []
    for t in transforms:
        a = t.augment_image(grid)
        if a.shape[2:] != (H, W) and a.shape[2:] != (W, H):
            raise ValueError("transform changes the image size")
        aug.append(_fit_map(a[0]))
        out_grid = torch.stack(torch.meshgrid(torch.arange(a.shape[2]), torch.arange(a.shape[3]),
                                              indexing="ij")).float()[None]
        deaug.append(_fit_map(t.deaugment_mask(out_grid)[0]))
    cache[(H, W)] = (aug, deaug)
    return aug, deaug


def is_fusable(transforms: Compose) -> bool:
    return all(isinstance(t, (HorizontalFlip, VerticalFlip, Rotate90)) for t in transforms.aug_transforms) \
        and 1 <= len(transforms) <= 8


class SegmentationTTAWrapper(nn.Module):
    """Drop-in for ``ttach.SegmentationTTAWrapper`` (tta.py:95-99,176-180)."""

    def __init__(self, model: nn.Module, transforms: Compose, merge_mode: str = "mean", output_mask_key=None):
        super().__init__()
        self.model = model
        self.transforms = transforms
        self.merge_mode = merge_mode
        self.output_key = output_mask_key

    def forward(self, image: torch.Tensor, *args):
        fused = getattr(self.model, "forward_tta", None)
        if fused is not None and self.merge_mode == "mean" and self.output_key is None and not args \
                and is_fusable(self.transforms) and image.shape[2] == image.shape[3]:
            return fused(image, self.transforms)
        merger = Merger(type=self.merge_mode, n=len(self.transforms))
        for transformer in self.transforms:
            augmented_output = self.model(transformer.augment_image(image), *args)
            if self.output_key is not None:
                augmented_output = augmented_output[self.output_key]
            merger.append(transformer.deaugment_mask(augmented_output))
        result = merger.result
        if self.output_key is not None:
            result = {self.output_key: result}
        return result
