"""Import the UNMODIFIED reference on CPU: from ``/root/reference`` where that exists (this container), else from the
byte-identical staged copy ``oracle/_ref/`` (``oracle/stage_reference.py``; what travels to the GPU box).

Test infrastructure: used by ``tests/golden/make_golden.py`` (fixture
generation) and by the ``needs_reference`` tests that pin ``oracle/`` against
the real reference, and by ``bench.py``'s CPU-baseline legs (the thing being
timed there, never the product).  ``/root/reference`` does not exist on the
GPU box: there this module resolves to the staged copy, or reports the
reference as unavailable.

The reference imports a few packages at module top that are absent from this
image and are not on the inference path: ``thop`` /
``torchinfo`` (profiling helpers, nets/Achelous.py:5-7) and ``timm`` (only
``DropPath`` = identity in eval, ``trunc_normal_`` = initialiser and
``register_model`` = decorator are touched by the EN/MV + GDF path).  They are
shimmed with inert stand-ins; no reference source is altered or copied.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    for root in (os.environ.get("ACHELOUS_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if root and os.path.isfile(os.path.join(root, "nets", "Achelous.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nets", "Achelous.py"))


def reference_kind() -> str:
    """'source tree' (/root/reference) or 'staged copy' (oracle/_ref) - reported next to every timing of the reference"""
    return "staged copy oracle/_ref" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED) else f"source tree {REFERENCE_ROOT}"


def _install_shims():
    import torch
    import torch.nn as nn

    def _mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    if "thop" not in sys.modules:
        thop = _mod("thop")
        thop.profile = lambda *a, **k: (0, 0)
        thop.clever_format = lambda *a, **k: ("0", "0")
    if "torchinfo" not in sys.modules:
        _mod("torchinfo").summary = lambda *a, **k: None
    try:
        import timm  # noqa: F401
        return
    except Exception:
        pass

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, *a, **k):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training or self.drop_prob == 0.0
            return x

    class SqueezeExcite(nn.Module):  # imported by out-of-scope backbones only
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, x):
            raise NotImplementedError

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def register_model(fn):
        return fn

    def _cfg(url="", **kwargs):
        return dict(url=url, **kwargs)

    timm = _mod("timm")
    timm.__path__ = []
    models = _mod("timm.models")
    models.__path__ = []
    layers = _mod("timm.models.layers")
    layers.__path__ = []
    helpers = _mod("timm.models.layers.helpers")
    registry = _mod("timm.models.registry")
    vit = _mod("timm.models.vision_transformer")
    data = _mod("timm.data")
    timm.models, timm.data = models, data
    models.layers, models.registry, models.vision_transformer = layers, registry, vit
    models.register_model = register_model
    layers.helpers = helpers
    for m in (layers,):
        m.DropPath, m.trunc_normal_, m.to_2tuple, m.SqueezeExcite = DropPath, trunc_normal_, to_2tuple, SqueezeExcite
    helpers.to_2tuple = to_2tuple
    registry.register_model = register_model
    vit._cfg, vit.trunc_normal_ = _cfg, trunc_normal_
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


_cache = {}


def load_reference():
    """Returns a namespace with the reference's Achelous / Achelous3T classes and
    utils_bbox free functions, imported from REFERENCE_ROOT."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    nets = importlib.import_module("nets.Achelous")
    bbox = importlib.import_module("utils.utils_bbox")
    ns = types.SimpleNamespace(Achelous=nets.Achelous, Achelous3T=nets.Achelous3T,
                               decode_outputs=bbox.decode_outputs,
                               non_max_suppression=bbox.non_max_suppression)
    _cache["ns"] = ns
    return ns
