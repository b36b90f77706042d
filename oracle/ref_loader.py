"""TEST INFRASTRUCTURE -- imports the reference's OWN hot-path files by path.

Only usable where ``/root/reference`` exists (the build container).  The reference
package cannot be imported normally (``archs/__init__.py:3`` pulls every architecture and
with them timm / smp / pytorch_toolbelt / catalyst, none of which are installed), so the
individual files are loaded inside synthetic packages after registering the third-party
restatements of ``oracle/shims.py`` in ``sys.modules``.

``load_tta()`` / ``load_ensemble()`` go one level up and execute the reference's DRIVER files
(src/main/tta.py, ensemble.py) unmodified, with its own dataset / transform / tiling / scoring code.

Used by tests/golden/make_golden.py (to generate the committed fixtures) and by
tests/test_oracle.py (to pin ``oracle/nets.py``, ``oracle/scoring.py`` and ``oracle/pipeline.py``
against the reference's own code).  Never imported by the product package.
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


def build_reference_module(ref, name, cfg):
    """The reference module behind a registry name, built the way ``archs.get_model(..., training=False)`` builds
    it (archs/__init__.py:101-130)."""
    cfg = dict(cfg)
    if name == "unetplusplusstar":
        return ref.unetplusplusstar.UnetPlusPlusStar(**cfg)
    if name == "unetplusplus_deepsup":
        cfg["deep_supervision"] = False          # archs/__init__.py:118-119 (training=False)
        return ref.deep_supunetplusplus.UnetPlusPlus(**cfg)
    if name == "Unet":
        return ref.smp.Unet(**cfg)
    raise KeyError(name)


def _install_inference_stubs():
    """Third-party packages the reference's inference drivers import and this image lacks, restated:
    ttach 0.0.3 (views + mean merge = oracle.nets.tta_mean_logits), the albumentations 1.0 classes / functions the
    drivers touch (cv2 underneath, as in albumentations), rasterio's windowed read (PIL decode + slice),
    catalyst.dl.utils.get_device (= cpu)."""
    import cv2
    import numpy as np
    import torch
    from . import nets

    class SegmentationTTAWrapper(torch.nn.Module):
        def __init__(self, model, transforms, merge_mode="mean"):
            super().__init__()
            assert merge_mode == "mean"
            self.model, self.kind = model, transforms

        def forward(self, x):
            return nets.tta_mean_logits(self.model, x, self.kind)

    _module("ttach", SegmentationTTAWrapper=SegmentationTTAWrapper,
            aliases=types.SimpleNamespace(d4_transform=lambda: "d4", flip_transform=lambda: "flip",
                                          hflip_transform=lambda: "hflip"))
    _module("pytorch_toolbelt.inference")
    _module("pytorch_toolbelt.inference.tiles", ImageSlicer=object, TileMerger=object)
    _module("pytorch_toolbelt.utils", fs=None, image_to_tensor=None)
    _module("pytorch_toolbelt.utils.torch_utils", to_numpy=None, image_to_tensor=None, tensor_from_rgb_image=None)
    _module("iglovikov_helper_functions")
    _module("iglovikov_helper_functions.utils")
    _module("iglovikov_helper_functions.utils.image_utils", pad=None)
    dl_utils = _module("catalyst.dl.utils", get_device=lambda: torch.device("cpu"))
    _module("catalyst.dl", utils=dl_utils)

    # ---- albumentations 1.0 (3P), the classes / functions this path touches
    def longest_max_size(img, max_size, interpolation):
        h, w = img.shape[:2]
        scale = max_size / float(max(h, w))
        if scale == 1.0:
            return img
        return cv2.resize(img, (int(round(w * scale)), int(round(h * scale))), interpolation=interpolation)

    def resize(img, height, width, interpolation=cv2.INTER_LINEAR):
        if img.shape[:2] == (height, width):
            return img
        return cv2.resize(img, (width, height), interpolation=interpolation)

    def center_crop(img, crop_height, crop_width):
        h, w = img.shape[:2]
        if h < crop_height or w < crop_width:
            raise ValueError("Requested crop size is larger than the image size")
        y1, x1 = (h - crop_height) // 2, (w - crop_width) // 2
        return img[y1:y1 + crop_height, x1:x1 + crop_width]

    class _Dual:
        def __call__(self, **d):
            d["image"] = self.apply(d["image"], cv2.INTER_LINEAR)
            if "mask" in d:
                d["mask"] = self.apply(d["mask"], cv2.INTER_NEAREST)
            return d

    class LongestMaxSize(_Dual):
        def __init__(self, max_size=1024, **_):
            self.max_size = max_size

        def apply(self, a, interpolation):
            return longest_max_size(a, self.max_size, interpolation)

    class Resize(_Dual):
        def __init__(self, height, width, **_):
            self.h, self.w = height, width

        def apply(self, a, interpolation):
            return resize(a, self.h, self.w, interpolation)

    class PadIfNeeded(_Dual):
        def __init__(self, min_height, min_width, border_mode=cv2.BORDER_CONSTANT, value=0, **_):
            assert border_mode == cv2.BORDER_CONSTANT and value == 0
            self.h, self.w = min_height, min_width

        def apply(self, a, interpolation):
            h, w = a.shape[:2]
            top = int((self.h - h) / 2.0) if h < self.h else 0
            left = int((self.w - w) / 2.0) if w < self.w else 0
            out = np.zeros((max(self.h, h), max(self.w, w)) + a.shape[2:], dtype=a.dtype)
            out[top:top + h, left:left + w] = a
            return out

    class Lambda:
        def __init__(self, image=None, **_):
            self.fn = image

        def __call__(self, **d):
            d["image"] = self.fn(d["image"])
            return d

    class ToTensorV2:
        def __call__(self, **d):
            d["image"] = torch.from_numpy(np.ascontiguousarray(d["image"].transpose(2, 0, 1)))
            if "mask" in d:
                d["mask"] = torch.from_numpy(np.ascontiguousarray(d["mask"]))
            return d

    class Compose:
        def __init__(self, transforms, **_):
            self.transforms = list(transforms)

        def __call__(self, **d):
            d = dict(d)
            for t in self.transforms:
                d = t(**d)
            return d

    alb = _module("albumentations", Compose=Compose, Lambda=Lambda, LongestMaxSize=LongestMaxSize,
                  PadIfNeeded=PadIfNeeded, Resize=Resize)
    _module("albumentations.pytorch", ToTensorV2=ToTensorV2)
    _module("albumentations.pytorch.transforms", ToTensorV2=ToTensorV2)
    alb.pytorch = sys.modules["albumentations.pytorch"]
    _module("albumentations.augmentations")
    _module("albumentations.augmentations.crops")
    _module("albumentations.augmentations.crops.functional", center_crop=center_crop)
    _module("albumentations.augmentations.geometric")
    _module("albumentations.augmentations.geometric.functional", longest_max_size=longest_max_size, resize=resize)
    _module("albumentations.augmentations.geometric.resize", RandomScale=object)

    # ---- rasterio 1.2 (3P): tta.py:196,201 open a JPEG and read RGB windows
    class Window:
        def __init__(self, rows, cols):
            self.rows, self.cols = rows, cols

        @classmethod
        def from_slices(cls, rows, cols):
            return cls(tuple(int(v) for v in rows), tuple(int(v) for v in cols))

    class _Dataset:
        def __init__(self, path):
            from PIL import Image
            self._a = np.asarray(Image.open(path).convert("RGB"))
            self.shape = self._a.shape[:2]

        def read(self, indexes, window=None):
            a = self._a if window is None else self._a[window.rows[0]:window.rows[1], window.cols[0]:window.cols[1]]
            return np.stack([a[..., i - 1] for i in indexes])

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    _module("rasterio", open=lambda path, *a, **k: _Dataset(path), Affine=lambda *a: a)
    _module("rasterio.windows", Window=Window)


def _reference_data_classes():
    """The reference's own src/data/lesion_dataset.py and data_transform.py, loaded by path."""
    data_dir = os.path.join(REFERENCE_ROOT, "src", "data")
    _module("refdatapkg")
    return (_load("refdatapkg", "lesion_dataset", os.path.join(data_dir, "lesion_dataset.py")),
            _load("refdatapkg", "data_transform", os.path.join(data_dir, "data_transform.py")))


def load_ensemble():
    """The reference's top-level ``ensemble.py`` executed UNMODIFIED, together with the reference's own
    ``src/data/lesion_dataset.py`` (TestSegmentation), ``src/data/data_transform.py`` (NormalTransform),
    ``base_utils.get_datapath`` / ``save_output`` and ``aucpr.py``.

    Third-party packages that are not installed are restated: ``ttach`` by
    ``oracle.nets.tta_mean_logits``, the five ``albumentations`` classes the path touches (Compose, Lambda,
    LongestMaxSize, PadIfNeeded, ToTensorV2) with cv2, ``catalyst.dl.utils.get_device`` = cpu.

    ``ensemble.py`` is stale against the rest of the reference in four places (listed in the product's
    ``ensemble.py`` docstring); they are bridged here by ADAPTERS around the reference's own functions, not by
    editing it: ``TestSegmentation(images, masks, transform=)`` -> ``TestSegmentation(images, False, masks, ...)``;
    ``get_preprocessing_fn(dataset_name=)`` -> ``grayscale=False``; ``get_auc(gts, preds, config)`` /
    ``plot_aucpr_curve(gts, preds, outdir, config)`` -> the ``(pred, gt, name)`` generator form of aucpr.py:17,45
    (first two thresholds returned).  Returns ``(module, captured)``: ``captured`` collects what the adapters saw
    (``preds``, ``gts``, ``auc``, ``thresholds``, ``masks`` by file name)."""
    import numpy as np

    ref = load()
    _install_inference_stubs()
    captured = {"masks": {}}

    lesion_dataset, data_transform = _reference_data_classes()

    class TestSegmentation(lesion_dataset.TestSegmentation):        # adapter: the call of ensemble.py:78
        def __init__(self, images, masks=None, transform=None):
            super().__init__(images, False, masks, transform=transform)

    TestSegmentation.__test__ = False                               # not a pytest class

    def _items(gts, preds):
        return [(np.asarray(p), np.asarray(g).reshape(np.asarray(p).shape), str(i)) for i, (g, p) in enumerate(zip(gts, preds))]

    def get_auc(gt_masks, tta_predictions, config):                 # adapter: ensemble.py:102 -> aucpr.py:17
        captured["preds"], captured["gts"] = list(tta_predictions), list(gt_masks)
        captured["auc"] = float(ref.aucpr.get_auc(_items(gt_masks, tta_predictions), config))
        return captured["auc"]

    def plot_aucpr_curve(gt_masks, tta_predictions, outdir, config):  # adapter: ensemble.py:105 -> aucpr.py:45
        th = ref.aucpr.plot_aucpr_curve(_items(gt_masks, tta_predictions), outdir, config)
        captured["thresholds"] = [float(t) for t in th]
        return th[0], th[1]

    def save_output(mask, out_path):                                # the reference's own writer, observed
        captured["masks"][os.path.basename(str(out_path))] = np.array(mask)
        return ref.base_utils.save_output(mask, out_path)

    archs = _module("src.main.archs",
                    get_model=lambda model_name, params, training=False: build_reference_module(ref, model_name, params),
                    get_preprocessing_fn=lambda dataset_name, grayscale=False: get_preprocessing_fn(dataset_name, grayscale))
    _module("src")
    _module("src.main", archs=archs)
    _module("src.main.aucpr", get_auc=get_auc, plot_aucpr_curve=plot_aucpr_curve)
    _module("src.main.util", get_datapath=ref.base_utils.get_datapath, save_output=save_output)
    _module("src.data", NormalTransform=data_transform.NormalTransform, TestSegmentation=TestSegmentation)
    _module("refroot")
    sys.modules.pop("refroot.ensemble", None)
    mod = _load("refroot", "ensemble", os.path.join(REFERENCE_ROOT, "ensemble.py"))
    return mod, captured


def load_tta(which="tta"):
    """The reference's inference drivers, ``src/main/tta.py`` (lesions; ``which="tta_vessel"``: its vessel twin
    ``src/main/tta_vessel.py``, scored on ROC) -- ``test_tta``, ``tta_patches`` --, executed
    UNMODIFIED together with the reference's own ``TestSegmentation`` / ``NormalTransform`` / ``base_utils``
    (``make_grid``, ``multigen``, ``get_datapath``, ``save_output``) / ``aucpr.py`` and
    ``archs.get_preprocessing_fn``; ``archs.get_model`` builds the reference's own modules
    (``build_reference_module``).  Third-party packages: see ``_install_inference_stubs``.

    Two things stand between that file and a CPU-only container and are handled OUTSIDE it: its loader asks for
    worker processes and pinned memory, and ``test_tta`` sends the batch ``.to('cuda')`` (tta.py:111); the module's
    ``DataLoader`` name is therefore bound to an in-process loader whose image batches ignore ``.to``.

    Returns ``(module, captured)``; ``captured[call]`` = the ``(pred, gt, name)`` items each scoring call received,
    ``captured["auc"]``, ``captured["thresholds"]``, ``captured["masks"]`` (file name -> array handed to
    ``save_output``)."""
    import numpy as np
    import torch

    ref = load()
    _install_inference_stubs()
    lesion_dataset, data_transform = _reference_data_classes()
    captured = {"masks": {}}
    bu = ref.base_utils

    def get_auc(generator, config):
        items = [(np.array(p), np.array(g), str(n)) for p, g, n in generator]
        captured["items"] = items
        captured["auc"] = float(ref.aucpr.get_auc(items, config))
        return captured["auc"]

    def plot_aucpr_curve(generator, exp_name, config):
        th = ref.aucpr.plot_aucpr_curve(list(generator), exp_name, config)
        captured["thresholds"] = [float(t) for t in th]
        return th

    def get_aucroc(generator, config):
        items = [(np.array(p), np.array(g), str(n)) for p, g, n in generator]
        captured["items"] = items
        captured["auc"] = float(ref.aucpr.get_aucroc(items, config))
        return captured["auc"]

    def plot_aucroc_curve(generator, exp_name, config):
        t = ref.aucpr.plot_aucroc_curve(list(generator), exp_name, config)
        captured["thresholds"] = [float(t)]
        return t

    def save_output(mask, out_path):
        captured["masks"][os.path.basename(str(out_path))] = np.array(mask)
        return bu.save_output(mask, out_path)

    archs = _module("refsrc.main.archs",
                    get_model=lambda model_name, params, training=False: build_reference_module(ref, model_name, params),
                    get_preprocessing_fn=get_preprocessing_fn)
    sys.modules["refsrc.main"].archs = archs
    # observed (not replaced) scoring / writer: the wrappers call the reference's functions
    _module("refsrc.main.aucpr", get_auc=get_auc, plot_aucpr_curve=plot_aucpr_curve, get_aucroc=get_aucroc,
            plot_aucroc_curve=plot_aucroc_curve)
    _module("refsrc.main.util", lesion_dict=bu.lesion_dict, get_datapath=bu.get_datapath, make_grid=bu.make_grid,
            multigen=bu.multigen, save_output=save_output)
    _module("refsrc.data", NormalTransform=data_transform.NormalTransform,
            TestSegmentation=lesion_dataset.TestSegmentation)
    assert which in ("tta", "tta_vessel")
    sys.modules.pop("refsrc.main." + which, None)
    mod = _load("refsrc.main", which, os.path.join(REFERENCE_ROOT, "src", "main", which + ".py"))

    class _HostTensor(torch.Tensor):
        def to(self, *args, **kwargs):           # tta.py:111 `.to('cuda')` on a CPU-only box
            return self.as_subclass(torch.Tensor)

    real_loader = torch.utils.data.DataLoader

    def loader(ds, **kw):
        kw.update(num_workers=0, pin_memory=False)
        for batch in real_loader(ds, **kw):
            batch["image"] = batch["image"].as_subclass(_HostTensor)
            yield batch

    class _Loader:
        def __init__(self, ds, **kw):
            self.ds, self.kw = ds, kw

        def __iter__(self):
            return loader(self.ds, **self.kw)

        def __len__(self):
            return -(-len(self.ds) // self.kw.get("batch_size", 1))

    mod.DataLoader = _Loader
    return mod, captured


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
