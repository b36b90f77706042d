"""TEST INFRASTRUCTURE -- imports the reference's OWN hot-path files by path.

Only usable where ``/root/reference`` exists (the build container).  The reference
package cannot be imported normally (``archs/__init__.py:3`` pulls every architecture and
with them timm / smp / pytorch_toolbelt / catalyst, none of which are installed), so the
individual files are loaded inside synthetic packages after registering the third-party
restatements of ``oracle/shims.py`` in ``sys.modules``.

Used by tests/golden/make_golden.py (to generate the committed fixtures) and by
tests/test_oracle.py (to pin ``oracle/nets.py`` and ``oracle/scoring.py`` against the
reference's own code).  Never imported by the product package.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EDS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "main", "archs"))


def _module(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package so dotted children resolve
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _install_third_party_stubs() -> None:
    from . import shims
    import torch

    # pytorch_toolbelt.modules.backbone.senet.se_resnet50
    _module("pytorch_toolbelt")
    _module("pytorch_toolbelt.modules")
    _module("pytorch_toolbelt.modules.backbone")
    _module("pytorch_toolbelt.modules.backbone.senet", se_resnet50=shims.se_resnet50)

    # segmentation_models_pytorch
    md = _module("segmentation_models_pytorch.base.modules", Attention=shims.Attention, Activation=shims.Activation,
                 Flatten=shims.Flatten, Conv2dReLU=shims.Conv2dReLU, SCSEModule=shims.SCSEModule)
    init = _module("segmentation_models_pytorch.base.initialization", initialize_decoder=shims.initialize_decoder,
                   initialize_head=shims.initialize_head)
    base = _module("segmentation_models_pytorch.base", modules=md, initialization=init,
                   SegmentationModel=shims.SegmentationModel, SegmentationHead=shims.SegmentationHead,
                   ClassificationHead=shims.ClassificationHead)
    enc = _module("segmentation_models_pytorch.encoders", get_encoder=shims.get_encoder)
    dec = _module("segmentation_models_pytorch.unetplusplus.decoder", UnetPlusPlusDecoder=object)
    upp = _module("segmentation_models_pytorch.unetplusplus", decoder=dec)
    _module("segmentation_models_pytorch", base=base, encoders=enc, unetplusplus=upp, Unet=shims.Unet)

    # timm.models.layers
    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    layers = _module("timm.models.layers", DropBlock2d=shims.DropBlock2d, DropPath=shims.DropPath,
                     to_2tuple=to_2tuple, trunc_normal_=torch.nn.init.trunc_normal_)
    models = _module("timm.models", layers=layers)
    _module("timm", models=models)

    # plotting / logging-only dependencies of aucpr.py and base_utils.py
    class _Figure:
        def add_shape(self, *a, **k): pass
        def update_yaxes(self, *a, **k): pass
        def update_xaxes(self, *a, **k): pass
        def write_image(self, *a, **k): pass

    _module("plotly")
    _module("plotly.express", area=lambda *a, **k: _Figure())
    sys.modules["plotly"].express = sys.modules["plotly.express"]
    _module("catalyst")
    _module("catalyst.utils")
    _module("catalyst.utils.distributed", get_distributed_env=lambda *a, **k: {},
            get_distributed_params=lambda *a, **k: {})
    _module("prettytable", PrettyTable=object)


def _load(pkg: str, name: str, path: str) -> types.ModuleType:
    full = f"{pkg}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


_loaded = {}


def load():
    """-> namespace with the reference modules: unetplusplusstar, axial_attention_v2,
    deep_supunetplusplus, aucpr, base_utils, stat_result, stat_result_vessel, smp (the shimmed package)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_third_party_stubs()
    main = os.path.join(REFERENCE_ROOT, "src", "main")
    archs_dir = os.path.join(main, "archs")
    pkg = _module("refarchs")
    pkg.__path__ = [archs_dir]
    # `.modules` is only needed for BottleBlock, which use_axial=True never instantiates
    _module("refarchs.modules", BottleBlock=object)
    for name in ("model_util", "axial_attention_v2", "unetplusplusstar", "deep_supunetplusplus"):
        _loaded[name] = _load("refarchs", name, os.path.join(archs_dir, name + ".py"))
    _module("refutil")
    _loaded["base_utils"] = _load("refutil", "base_utils", os.path.join(main, "util", "base_utils.py"))
    refmain = _module("refmain")
    refmain.__path__ = [main]
    _module("refmain.util", lesion_dict=_loaded["base_utils"].lesion_dict)
    _loaded["aucpr"] = _load("refmain", "aucpr", os.path.join(main, "aucpr.py"))
    _loaded["smp"] = sys.modules["segmentation_models_pytorch"]
    # stat_result.py does `from .util import lesion_dict` and `from ..data import NormalTransform` (unused)
    refsrc = _module("refsrc")
    refsrc.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
    refsrc_main = _module("refsrc.main")
    refsrc_main.__path__ = [main]
    _module("refsrc.main.util", lesion_dict=_loaded["base_utils"].lesion_dict)
    _module("refsrc.data", NormalTransform=object)
    _loaded["stat_result"] = _load("refsrc.main", "stat_result", os.path.join(main, "stat_result.py"))
    _loaded["stat_result_vessel"] = _load("refsrc.main", "stat_result_vessel", os.path.join(main, "stat_result_vessel.py"))
    return types.SimpleNamespace(**_loaded)


def load_pad_img():
    """The reference's vessel padding script (src/data/augment_vessel/pad_img.py).  It imports scikit-image,
    which is not installed: ``skimage.io.imread(path, as_gray=True)`` is stubbed by a PIL read that returns
    what scikit-image returns for the single-channel 8-bit label files of DRIVE / CHASEDB1 (the 2-D uint8
    array, unchanged; as_gray only converts multi-channel images)."""
    if "refdata.pad_img" in sys.modules:
        return sys.modules["refdata.pad_img"]
    import numpy as np
    from PIL import Image

    def imread(path, as_gray=False):
        im = Image.open(path)
        if as_gray and im.mode not in ("L", "P", "1", "I;16"):
            raise NotImplementedError("the skimage stub only covers single-channel label files")
        return np.asarray(im.convert("L")) if as_gray else np.asarray(im)

    io = _module("skimage.io", imread=imread)
    tr = _module("skimage.transform")
    _module("skimage", io=io, transform=tr)
    _module("refdata")
    return _load("refdata", "pad_img", os.path.join(REFERENCE_ROOT, "src", "data", "augment_vessel", "pad_img.py"))


def get_preprocessing_fn(dataset_name, grayscale=False):
    """archs/__init__.py cannot be imported (it imports every architecture); its
    get_preprocessing_fn (lines 61-99) is executed from source text instead."""
    import numpy as np
    src = open(os.path.join(REFERENCE_ROOT, "src", "main", "archs", "__init__.py")).read()
    start = src.index("def get_preprocessing_fn")
    end = src.index("def list_models")
    ns = {"np": np}
    exec(compile(src[start:end], "ref_archs_init_fragment", "exec"), ns)
    return ns["get_preprocessing_fn"](dataset_name, grayscale)


def registry_names():
    """Keys of MODEL_REGISTRY (archs/__init__.py:7-59) parsed from source."""
    import re
    src = open(os.path.join(REFERENCE_ROOT, "src", "main", "archs", "__init__.py")).read()
    body = src[src.index("MODEL_REGISTRY = {"):src.index("def get_preprocessing_fn")]
    return re.findall(r'^\s*"([^"]+)"\s*:', body, flags=re.M)
