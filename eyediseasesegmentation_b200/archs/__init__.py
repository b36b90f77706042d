"""Mirror of the reference's model registry (src/main/archs/__init__.py:5-120) for the
inference-and-scoring hot path.

``get_model(model_name, params, training)`` / ``MODEL_REGISTRY`` / ``list_models`` /
``get_preprocessing_fn`` keep the reference's names, argument meaning and error behaviour.
The three networks on the path (``unetplusplusstar``, ``unetplusplus_deepsup`` and smp's
``Unet``) are constructed as :class:`B200SegModel`, an ``nn.Module`` whose parameters carry
the reference's ``state_dict`` keys and whose ``forward`` runs on the sm_100a kernels.  The
remaining registry names are kept (so ``list_models()`` agrees with the reference) but are
outside the B200 hot path and raise ``NotImplementedError`` when constructed (SURVEY.md 2,
row 11).
"""
from __future__ import annotations

import os
import threading
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import spec as _spec
from .engine import Engine

__all__ = ["list_models", "get_model", "get_preprocessing_fn", "MODEL_REGISTRY", "B200SegModel", "Unet"]


_GRAPH_POOLS: Dict[str, object] = {}


def _graphs_enabled() -> bool:
    return os.environ.get("EDS_CUDA_GRAPHS", "1") != "0"


def _graph_pool(device):
    """One memory pool per device shared by every captured forward: the graphs are replayed one at a time
    on one stream, so the intermediates of one graph can live in the space another one uses."""
    key = str(device)
    if key not in _GRAPH_POOLS:
        _GRAPH_POOLS[key] = torch.cuda.graph_pool_handle()
    return _GRAPH_POOLS[key]


class B200SegModel(nn.Module):
    """Segmentation network whose forward pass is hand-written CUDA for sm_100a.

    Call contract: ``[B,3,H,W] float32`` on a CUDA device -> logits
    ``[B,classes,H,W] float32``; supports ``load_state_dict`` with reference checkpoints,
    ``.eval()``, ``.to(device)``, wrapping by ``SegmentationTTAWrapper`` / ``nn.DataParallel``.
    ``precision`` is ``"bf16"`` (tcgen05 tensor cores, default) or ``"fp32"`` (parity mode;
    also selectable with EDS_PRECISION=fp32).

    Devices: the prepared (BN-folded, repacked) weights and the captured CUDA graphs are kept PER DEVICE and a
    forward runs with the input's device made current, so one process may drive several GPUs
    (``nn.DataParallel`` replicas share this state through their copied ``__dict__`` and each picks the
    entry of the device its input chunk lives on).  One process per GPU (torchrun) remains the fast path.
    """

    def __init__(self, arch: str, cfg: dict, seed: Optional[int] = None):
        super().__init__()
        self.arch = arch
        self.cfg = dict(cfg)
        gen = torch.Generator()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())  # follows torch.manual_seed like nn init
        gen.manual_seed(seed)
        s = _spec.ParamSpec(gen)
        classes = int(cfg.get("classes", 1))
        att = cfg.get("decoder_attention_type", None)
        use_bn = cfg.get("decoder_use_batchnorm", True)
        if use_bn == "inplace":
            raise RuntimeError("In order to use `use_batchnorm='inplace'` inplace_abn package must be installed.")
        if cfg.get("activation", None) is not None:
            raise NotImplementedError("the B200 path returns logits (activation=None), as every reference config does")
        if tuple(cfg.get("decoder_channels", _spec.DECODER_CHANNELS)) != _spec.DECODER_CHANNELS:
            raise NotImplementedError("decoder_channels other than (256,128,64,32,16) are outside the hot path")
        if int(cfg.get("encoder_depth", 5)) != 5 or int(cfg.get("in_channels", 3)) != 3:
            raise NotImplementedError("encoder_depth=5 and in_channels=3 are the only configurations on the hot path")
        if arch == "unetplusplusstar":
            enc = cfg["encoder_name"]
            if enc not in ("BoTSER50_Axial_Imagenet", "BoTSER50_Axial_Imagenet_2", "BoTSER50_Axial_Imagenet_3",
                           "BoTSER50_Axial_scratch"):
                raise KeyError(f"encoder_name {enc!r}: only the axial (MHSA) BoTSER50 encoders are on the B200 path")
            _spec.build_unetplusplusstar(s, int(cfg.get("base_dim", 32)), att, classes, use_bn)
            self.name = f"unetplusplus-{enc}"
        elif arch == "unetplusplus_deepsup":
            _spec.build_smp_style(s, cfg.get("encoder_name", "resnet34"), "unetplusplus", att, classes, use_bn,
                                  deep_heads=True)
            self.name = "unetplusplus-{}".format(cfg.get("encoder_name", "resnet34"))
        elif arch == "Unet":
            _spec.build_smp_style(s, cfg.get("encoder_name", "resnet34"), "unet", att, classes, use_bn)
            self.name = "u-{}".format(cfg.get("encoder_name", "resnet34"))
        else:
            raise KeyError(arch)
        _spec.materialise(self, s)
        self.deep_supervision = False
        self.precision = os.environ.get("EDS_PRECISION", "bf16")
        # shared (by reference) with nn.DataParallel replicas: {(device, precision): Engine}, {key: graph entry},
        # the module that owns the parameters, and a lock for concurrent replica threads
        self._engines: Dict[tuple, Engine] = {}
        self._graphs = {}
        self._owner = [self]
        self._lock = threading.Lock()

    # ---- any change to the parameters invalidates the prepared (folded) weights and captured graphs
    def _invalidate(self):
        self._engines.clear()
        self._graphs.clear()

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engines", "_graphs", "_owner", "_lock"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._engines, new._graphs, new._owner, new._lock = {}, {}, [new], threading.Lock()
        return new

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("B200SegModel is inference-only (training is outside the hot path)")
        return super().train(False)

    def engine(self, device: Optional[torch.device] = None) -> Engine:
        """Prepared weights on ``device`` (default: where the parameters live), built on first use."""
        owner = self._owner[0]
        dev = torch.device(device) if device is not None else next(owner.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("B200SegModel runs only on a CUDA (sm_100a) device: call .to('cuda') first; "
                               "there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        key = (str(dev), self.precision)
        eng = self._engines.get(key)
        if eng is None:
            with self._lock:
                eng = self._engines.get(key)
                if eng is None:
                    with torch.no_grad(), torch.cuda.device(dev):
                        eng = Engine(self.arch, self.cfg, owner.state_dict(), dev, self.precision)
                    self._engines[key] = eng
        return eng

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("the B200 networks only run on a CUDA device (no CPU fallback); got a CPU tensor")
        with torch.cuda.device(x.device):
            return self.engine(x.device).run(x)

    @torch.no_grad()
    def forward_tta(self, x: torch.Tensor, transforms, apply_sigmoid: bool = False, merge: bool = True) -> torch.Tensor:
        """All TTA views in one batched pass: views folded into the stem loader, merged by
        ``eds_tta_merge``.  Returns the mean logits ``[B,1,H,W]`` (what
        ``SegmentationTTAWrapper.forward`` returns) or the probabilities if apply_sigmoid.

        The ~250 kernel launches of a pass are captured into a CUDA graph the second time a given
        (input shape, views, precision) is seen and replayed afterwards: at the tile-batch sizes of
        the drivers a third of the launches are shorter than their host-side issue cost, and the
        graph removes those gaps (set EDS_CUDA_GRAPHS=0 to run eagerly).

        ``merge=False`` returns the per-view logits ``[V,B,H,W]`` (view-major, still augmented) for the fused
        blend kernel ``eds_tta_blend_x2_f32`` instead of merging them here."""
        from .. import ttach_compat as tta
        from .. import kernels as K
        B, _, H, W = x.shape
        aug, deaug = tta.view_maps(transforms, H, W)
        if not x.is_cuda:
            raise RuntimeError("the B200 networks only run on a CUDA device (no CPU fallback); got a CPU tensor")
        with torch.cuda.device(x.device):
            return self._forward_tta_on_device(x, aug, deaug, apply_sigmoid if merge else None)

    def _forward_tta_on_device(self, x, aug, deaug, apply_sigmoid):
        # apply_sigmoid None = no merge: per-view logits
        from .. import kernels as K
        use_graph = (_graphs_enabled() and K.CONV_TRACE is None and K.KERNEL_TRACE is None and not self.engine(x.device).keep_features)
        if not use_graph:
            return self._forward_tta_eager(x, aug, deaug, apply_sigmoid)
        key = (tuple(x.shape), str(x.device), self.precision, tuple(map(tuple, aug)), apply_sigmoid)
        entry = self._graphs.get(key)
        if entry is None:                       # first sight: eager (also warms every lazy init)
            self._graphs[key] = "warm"
            return self._forward_tta_eager(x, aug, deaug, apply_sigmoid)
        if entry == "warm":
            static_x = x.detach().float().contiguous().clone()
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            launches0 = K.LAUNCHES[0]
            with torch.cuda.graph(graph, pool=_graph_pool(x.device)):
                static_y = self._forward_tta_eager(static_x, aug, deaug, apply_sigmoid)
            n_kernels = K.LAUNCHES[0] - launches0          # kernel nodes of the graph
            K.LAUNCHES[0] = launches0                       # capture launched nothing
            entry = self._graphs[key] = (graph, static_x, static_y, n_kernels)
        graph, static_x, static_y, n_kernels = entry
        static_x.copy_(x)
        graph.replay()
        K.LAUNCHES[0] += n_kernels
        # per-view logits (merge=False) are consumed by the blend kernel on this stream before the next replay:
        # the graph's own output buffer is handed out (valid until the next call); merged maps are copied
        return static_y if apply_sigmoid is None else static_y.clone()

    def _forward_tta_eager(self, x, aug, deaug, apply_sigmoid):
        from .. import kernels as K
        B, _, H, W = x.shape
        logits = self.engine(x.device).run(x, aug)            # [V*B, classes, H, W]
        if logits.shape[1] != 1 or H != W:
            raise NotImplementedError("fused TTA merge handles square single-class maps")
        if apply_sigmoid is None:
            return logits.view(len(aug), B, H, W)
        merged = K.tta_merge(logits.view(len(aug), B, H, W), deaug, apply_sigmoid)
        return merged.view(B, 1, H, W)

    def get_num_parameters(self):
        total = int(sum(p.numel() for p in self.parameters()))
        return total, total


def _star(**params):
    params.pop("drop_block_prob", None)     # DropBlock is the identity in eval mode
    params.pop("deep_supervision", None)    # forced False by get_model(training=False)
    params.pop("clf_head", None)
    return B200SegModel("unetplusplusstar", params)


def _unetplusplus_deepsup(**params):
    params.pop("deep_supervision", None)
    if params.pop("encoder_weights", None) is not None:
        raise RuntimeError("pretrained encoder weights cannot be downloaded here; load a checkpoint instead")
    params.pop("aux_params", None)
    return B200SegModel("unetplusplus_deepsup", params)


def Unet(**params):
    """smp-style constructor reached through ``getattr(smp, model_name)(**params)`` (tta.py:40-46)."""
    if params.pop("encoder_weights", None) is not None:
        raise RuntimeError("pretrained encoder weights cannot be downloaded here; load a checkpoint instead")
    params.pop("aux_params", None)
    return B200SegModel("Unet", params)


def _outside_hot_path(name):
    def ctor(*a, **k):
        raise NotImplementedError(
            f"model {name!r} is in the reference registry but outside the B200 hot path "
            "(unetplusplusstar, unetplusplus_deepsup, smp Unet); run it with the reference's PyTorch code")
    ctor.__name__ = name
    return ctor


_REFERENCE_NAMES = [
    "resnet50_attunet", "seresnet50_attunet", "efficientnetb2_attunet", "mobilenetv3_attunet", "swin_tiny_attunet",
    "swin_small_attunet", "hrnet18", "hrnet34", "hrnet48", "resnet50_doubleunet", "efficientnetb2_doubleunet",
    "mobilenetv3_doubleunet", "vgg_doubleunet", "unet_resnext50_ssl", "rrcnn_unet", "sa_unet", "hed_unet",
    "hed_resunet", "hed_denseunet", "resnet18_unet32", "resnet34_unet32", "resnet50_unet32", "b4_unet32",
    "b4_effunet32", "b2_effunet32", "b2_fpn_cat", "seresnext50_fpncat128", "resnet34_fpncat128",
    "resnet152_fpncat256", "transunet_r50", "transunet_b16", "unetplusplusstar", "LeeJunHyun_impl_att",
    "LeeJunHyun_impl_R2U_Net", "LeeJunHyun_impl_R2AttU_Net", "Unet3Plus_Base", "Unet3Plus_DS", "axialatt_unet",
    "gated", "medt", "logo", "axialattwopo_unet", "dcunet", "resunetplusplus", "unetplusplus_deepsup",
    "hubmap_kaggle", "deeplabv3plus_deepsup", "TransUnet_V2", "SegFormerStar", "SwinformerStar",
]

MODEL_REGISTRY: Dict[str, object] = {n: _outside_hot_path(n) for n in _REFERENCE_NAMES}
MODEL_REGISTRY["unetplusplusstar"] = _star
MODEL_REGISTRY["unetplusplus_deepsup"] = _unetplusplus_deepsup


def get_preprocessing_fn(dataset_name: str, grayscale: bool):
    """Per-dataset mean/std and the float64 normaliser (archs/__init__.py:61-99)."""
    table = {
        "IDRiD": ([0.44976714, 0.2186806, 0.06459363], [0.33224553, 0.17116262, 0.086509705]),
        "FGADR": ([0.4554011, 0.2591345, 0.13285689], [0.28593522, 0.185085, 0.13528904]),
        "DDR": ([0.31897065, 0.19916488, 0.08322998], [0.32040685, 0.20822203, 0.114768185]),
        "DRIVE": ([0.49742976, 0.27066445, 0.16217253], [0.34794736, 0.18998094, 0.1084089]),
        "HRF": ([0.6273858, 0.20169912, 0.10424815], [0.2866019, 0.11408445, 0.060513902]),
        "CHASEDB1": ([0.4527923, 0.16221291, 0.028265305], [0.36041078, 0.14167951, 0.036878455]),
    }
    mean, std = table.get(dataset_name, table["IDRiD"])
    mean, std = list(mean), list(std)
    if grayscale:
        mean = mean[0] * 0.2989 + mean[1] * 0.5870 + mean[2] * 0.1140
        std = std[0] * 0.2989 + std[1] * 0.5870 + std[2] * 0.1140

    def preprocessing(x, mean=mean, std=std, **kwargs):
        x = x / 255.0
        if mean is not None:
            x = x - np.array(mean)
        if std is not None:
            x = x / np.array(std)
        return x

    return preprocessing, mean, std


def list_models():
    return list(MODEL_REGISTRY.keys())


def get_model(model_name: str, params=None, training=True) -> nn.Module:
    try:
        model_fn = MODEL_REGISTRY[model_name]
    except KeyError:
        raise KeyError(f"Cannot found {model_name}, available options are {list(MODEL_REGISTRY.keys())}")
    if params is None:
        return model_fn()
    if not training:  # same in-place overrides as archs/__init__.py:111-119
        if params.get("clfhead", None) is not None:
            params["clfhead"] = False
        if params.get("pretrained", None) is not None:
            params["pretrained"] = False
        if params.get("encoder_weights", None) is not None:
            params["encoder_weights"] = None
        if params.get("deep_supervision", None) is not None:
            params["deep_supervision"] = False
    else:
        raise NotImplementedError("training=True is outside the B200 inference hot path")
    return model_fn(**params)
