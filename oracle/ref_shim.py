"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference from /root/reference.

The reference (gaopinghai/BDE2VID) is pure Python/PyTorch but depends on packages that are
not in this image (mmengine, mmcv, timm, h5py, matplotlib, LPIPS internals).  None of them
do arithmetic on the hot path: they supply base classes, registries and init helpers.  This
module installs minimal stand-ins in ``sys.modules`` so the reference's own model and
voxeliser code can be imported and run on CPU, which is how the golden vectors under
``tests/golden`` are produced (see ``oracle/make_golden.py``) and how the CPU restatement in
``oracle/oracle_torch.py`` is pinned.

``/root/reference`` exists only in the build container; nothing that runs on the GPU box
imports this file.  Nothing in ``bde2vid_b200`` may import anything from ``oracle/``.

Stub list follows SURVEY.md section 8(c).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("BDE2VID_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model", "BDE2VID"))


class _Registry:
    """dict registry: ``register_module()`` decorator and ``build(cfg)`` = pop 'type', call."""

    def __init__(self, name="registry", **kwargs):
        self.name = name
        self._mods = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._mods[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        return self._mods[typ](**cfg)

    def get(self, key):
        return self._mods.get(key)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Install the stub modules and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    MODELS = _Registry("model")
    METRICS = _Registry("metric")

    class _Config(dict):
        @staticmethod
        def fromstring(s, ext):
            ns = {}
            exec(s, ns)
            cfg = _Config({k: v for k, v in ns.items() if not k.startswith("__")})
            return cfg

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    _mod("mmengine", Registry=_Registry)
    _mod("mmengine.model", BaseModel=_BaseModule, BaseModule=_BaseModule)
    _mod("mmengine.registry", MODELS=MODELS, METRICS=METRICS)
    _mod("mmengine.config", Config=_Config)
    _mod("mmengine.evaluator", BaseMetric=object)

    def _kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
        if distribution == "uniform":
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def _constant_init(module, val, bias=0):
        if getattr(module, "weight", None) is not None:
            nn.init.constant_(module.weight, val)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    class _Deform(nn.Module):  # subclassed at import time by the reference, never instantiated here
        def __init__(self, *a, **k):
            super().__init__()

    _mod("mmcv")
    _mod("mmcv.cnn", constant_init=_constant_init, kaiming_init=_kaiming_init)
    _mod("mmcv.utils")
    _mod("mmcv.utils.parrots_wrapper", _BatchNorm=nn.modules.batchnorm._BatchNorm)
    _mod("mmcv.ops", DeformConv2dPack=_Deform, DeformConv2d=_Deform)
    _mod("mmcv.ops.deform_conv", deform_conv2d=None)

    class _DropPath(nn.Module):  # identity in eval mode, which is the only mode the oracle uses
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training, "oracle runs the reference in eval() only"
            return x

    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=_DropPath, trunc_normal_=nn.init.trunc_normal_)

    _mod("h5py")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            _mod("matplotlib")
            _mod("matplotlib.pyplot")
    _mod("LPIPS")
    _mod("LPIPS.util", util=types.SimpleNamespace())
    _mod("LPIPS.models", pretrained_networks=types.SimpleNamespace(), dist_model=types.SimpleNamespace())
    _mod("LPIPS.models.pretrained_networks")
    _mod("LPIPS.models.dist_model")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class cpu_mode:
    """Context manager: make ``Tensor.cuda`` a no-op (the generator calls it unconditionally,
    bde2vid_cross_scale_propogation_V5.py:153,166) so the reference runs on CPU."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False


def reference_modules():
    """Return a namespace with the reference classes/functions on the hot path."""
    install()
    from model.BDE2VID.bde2vid import BDE2VID
    from model.e2vid.model import E2VIDRecurrent
    from events_contrast_maximization.utils.event_utils import events_to_voxel_torch
    from utils_func.inference_utils import Croper
    return types.SimpleNamespace(BDE2VID=BDE2VID, E2VIDRecurrent=E2VIDRecurrent,
                                 events_to_voxel_torch=events_to_voxel_torch, Croper=Croper)
