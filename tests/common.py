"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (phi, backbone, weight_seed, input_seed); must match tests/golden/make_golden.py
GOLDEN_CONFIGS = {
    "en_gdf_pn_s0": ("S0", "en", 2, 11),
    "en_gdf_pn_s2": ("S2", "en", 0, 12),
    "mv_gdf_pn_s0": ("S0", "mv", 0, 13),
    "en_cdf_pn_s0": ("S0", "en", 3, 14),   # CSP-Dual-FPN neck
    "ev_gdf_pn_s0": ("S0", "ev", 4, 15),   # EdgeViT backbone
    "ef_gdf_pn_s0": ("S0", "ef", 5, 16),   # EfficientFormerV2 backbone
}
GOLDEN_NECK = {"en_cdf_pn_s0": "cdf"}      # every other config uses the Ghost-Dual-FPN


def neck_of(name):
    return GOLDEN_NECK.get(name, "gdf")

MODEL_KW = dict(num_det=7, num_seg=9, resolution=320, neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8,
                nano_head=True, spp=True)
WH_BIAS = 1.3

# North-star tolerance: 1e-3 relative (BASELINE.json).  "Relative" is measured against the
# tensor's own scale: max|a-b| / max|b|.
REL_TOL = 1e-3


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def load_keys(name):
    with open(os.path.join(GOLDEN_DIR, name + ".keys.json")) as f:
        spec = json.load(f)
    return {k: (tuple(v[0]), getattr(torch, v[1])) for k, v in spec.items()}


def rel_err(a, b):
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    a = a.reshape(b.shape)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def summarize(t, n=512):
    t = torch.as_tensor(t).detach().float().cpu().contiguous().flatten()
    step = max(1, t.numel() // n)
    return np.concatenate([[t.mean().item(), t.std().item(), t.abs().max().item(), float(step)],
                           t[::step][:n].numpy()]).astype(np.float32)


def summary_rel_err(s_mine, s_gold):
    """Compares the strided sample of two `summarize` vectors, relative to the golden absmax."""
    scale = max(float(s_gold[2]), 1e-30)
    return float(np.abs(s_mine[4:] - s_gold[4:]).max() / scale)


def argmax_mismatch(logits, gold_argmax, margin=1e-4):
    """Fraction of pixels whose argmax differs from the golden one, (a) over all pixels and
    (b) over pixels whose top-2 margin exceeds `margin` * max|logit| (ReLU-ed logits tie often,
    SURVEY.md §0.5; ties and near-ties flip with any change of summation order)."""
    logits = torch.as_tensor(logits).float().cpu()
    gold = torch.as_tensor(gold_argmax).long()
    am = logits.argmax(1)
    top2 = logits.topk(2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > margin * logits.abs().max()
    diff = am != gold
    return diff.float().mean().item(), (diff & safe).float().sum().item(), safe.float().mean().item()
