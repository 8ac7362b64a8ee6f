"""Seeded, de-vacuated parameter fill for parity tests and benchmarks.

The reference's default initialisation makes a parity check vacuous
(SURVEY.md §0.4): EdgeNeXt layer-scales start at 1e-6
(conv_encoder.py:15, sdta_encoder.py:27,35), the DCN offset/modulator convs at
exactly 0 (dcn.py:29-30,39-40), ShuffleAttention gates at 0
(shuffle_attention.py:16-19) and BN running stats at (0, 1).  This module fills a
state-dict *by key name* so that every parameter matters, activations stay O(1)
through the depth of the network, and the values depend only on
``(seed, key, shape)`` - not on module construction order - so the reference,
the oracle and the CUDA module all get identical tensors on any machine.
"""
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(shape, g, mean=0.0, std=1.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def _uniform(shape, g, lo, hi):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def fill_state_dict(spec, seed: int = 0):
    """spec: mapping key -> tensor (or (shape, dtype)); returns a new dict of CPU tensors."""
    keys = list(spec.keys())
    shapes = {}
    dtypes = {}
    for k in keys:
        v = spec[k]
        if isinstance(v, torch.Tensor):
            shapes[k], dtypes[k] = tuple(v.shape), v.dtype
        else:
            shapes[k], dtypes[k] = tuple(v[0]), v[1]
    bn_prefixes = {k[: -len("running_mean")] for k in keys if k.endswith("running_mean")}
    out = {}
    for k in keys:
        shape, g = shapes[k], _gen(seed, k)
        leaf = k.rsplit(".", 1)[-1]
        prefix = k[: len(k) - len(leaf)]
        if leaf == "attention_bias_idxs":
            # EfficientFormerV2 relative-position index table (ImageEncoder.py:104-120,245-264): a function of the two grid sizes
            import itertools
            import math
            n_q, n_k = shape
            r_q, r_k = math.isqrt(n_q), math.isqrt(n_k)
            step = math.ceil(r_k / r_q)
            offsets, idxs = {}, []
            for p1 in itertools.product(range(r_q), range(r_q)):
                for p2 in itertools.product(range(r_k), range(r_k)):
                    off = (abs(p1[0] * step - p2[0]), abs(p1[1] * step - p2[1]))
                    if off not in offsets:
                        offsets[off] = len(offsets)
                    idxs.append(offsets[off])
            t = torch.tensor(idxs, dtype=torch.int64).view(n_q, n_k)
        elif leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=dtypes[k])
        elif leaf == "running_mean":
            t = _normal(shape, g, 0.0, 0.1)
        elif leaf == "running_var":
            t = _uniform(shape, g, 0.5, 1.5)
        elif prefix in bn_prefixes and leaf == "weight":
            t = _uniform(shape, g, 0.6, 1.4)
        elif prefix in bn_prefixes and leaf == "bias" and "seg_head" in k:
            t = _uniform(shape, g, 0.05, 0.3)  # keep part of the ReLU-ed seg logits > 0
        elif prefix in bn_prefixes and leaf == "bias":
            t = _normal(shape, g, 0.0, 0.1)
        elif leaf in ("gamma", "gamma_xca"):
            t = _uniform(shape, g, 0.5, 1.5)
        elif leaf in ("layer_scale_1", "layer_scale_2"):
            t = _uniform(shape, g, 0.1, 0.3)   # EfficientFormerV2: 14 residual blocks with BatchNorm-ed (unit-scale) branches
        elif leaf == "temperature":
            t = _uniform(shape, g, 0.5, 2.0)
        elif leaf in ("cweight", "sweight"):
            t = _normal(shape, g, 0.0, 1.0)
        elif leaf in ("cbias", "sbias"):
            t = _normal(shape, g, 1.0, 0.5)
        elif leaf == "bias" and "talking_head" in k:
            t = _normal(shape, g, 0.0, 0.002)  # added to every one of N_k attention entries of a row: keep the row sums O(1)
        elif leaf == "bias" and "offset_conv" in k:
            t = _uniform(shape, g, -1.0, 1.0)  # leave the integer sampling grid
        elif leaf == "bias":
            t = _normal(shape, g, 0.0, 0.1)
        elif leaf == "weight" and len(shape) == 1:
            t = _uniform(shape, g, 0.6, 1.4)  # LayerNorm / GroupNorm affine
        elif leaf == "weight":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 4.0 if "seg_head.primary_conv" in k else 1.0  # spatially varied seg argmax
            t = _normal(shape, g, 0.0, gain / max(fan_in, 1) ** 0.5)
        else:
            t = _normal(shape, g, 0.0, 0.1)
        out[k] = t.to(dtypes[k]).contiguous()
    return out
