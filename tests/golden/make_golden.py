"""Generate the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py shims) on CPU.

    python tests/golden/make_golden.py

Writes, per config, ``tests/golden/<cfg>.npz`` (final outputs, per-block taps
captured with forward hooks on the reference's own modules, decode + NMS
results) and ``tests/golden/<cfg>.keys.json`` (the strict state-dict contract).
The reference cannot travel to the GPU box; these files can.  Weights and
inputs are regenerated from seeds (achelous_b200/weights.py, synthetic.py), so
only outputs are stored.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    "en_gdf_pn_s0": dict(phi="S0", backbone="en", weight_seed=2, input_seed=11),
    "en_gdf_pn_s2": dict(phi="S2", backbone="en", weight_seed=0, input_seed=12),
    "mv_gdf_pn_s0": dict(phi="S0", backbone="mv", weight_seed=0, input_seed=13),
    "en_cdf_pn_s0": dict(phi="S0", backbone="en", weight_seed=3, input_seed=14, neck="cdf"),
    "ev_gdf_pn_s0": dict(phi="S0", backbone="ev", weight_seed=4, input_seed=15),
    "ef_gdf_pn_s0": dict(phi="S0", backbone="ef", weight_seed=5, input_seed=16),
}
MODEL_KW = dict(num_det=7, num_seg=9, resolution=320, pc_seg="pn", pc_channels=5, pc_classes=8,
                nano_head=True, spp=True)
WH_BIAS = 1.3  # widens boxes to ~3.7 cells so that NMS actually suppresses
TARGET_CANDIDATES = 120  # obj logits are shifted so ~this many anchors/img pass conf >= 0.35


def calibrate_obj_bias(det, conf=0.35, target=TARGET_CANDIDATES):
    """Bisect a scalar added to the obj logits until ~target anchors per image pass the threshold."""
    flat = torch.cat([d.flatten(2) for d in det], 2)  # (B, 12, A)
    cls = torch.sigmoid(flat[:, 5:]).max(1)[0]
    lo, hi = -20.0, 20.0
    for _ in range(40):
        mid = (lo + hi) / 2
        n = ((torch.sigmoid(flat[:, 4] + mid) * cls) >= conf).float().sum(1).mean().item()
        lo, hi = (mid, hi) if n < target else (lo, mid)
    return round((lo + hi) / 2, 3)

# reference module path -> oracle tap name (oracle/functional.py)
def tap_map(phi, backbone):
    m = {}
    fpn = "image_radar_encoder.fpn."
    if backbone == "en":
        depths = {"S0": [2, 2, 6, 2], "S2": [3, 3, 9, 3]}[phi]
        for i in range(4):
            m[f"{fpn}backbone.downsample_layers.{i}"] = f"backbone.down{i}"
            for j in range(depths[i]):
                m[f"{fpn}backbone.stages.{i}.{j}"] = f"backbone.stage{i}.{j}"
    if backbone == "ef":
        depth = {"S0": [2, 2, 6, 4], "S1": [3, 3, 9, 6], "S2": [4, 4, 12, 8]}[phi]
        for i in range(4):
            for j in range(depth[i]):
                m[f"{fpn}backbone.network.{2 * i}.{j}"] = f"backbone.stage{i}.{j}"
    if backbone == "ev":
        depth = {"S0": [1, 1, 3, 2], "S1": [1, 1, 3, 1], "S2": [1, 2, 5, 3]}[phi]
        for i in range(4):
            for j in range(depth[i]):
                m[f"{fpn}backbone.blocks{i + 1}.{j}"] = f"backbone.stage{i}.{j}"
    m[f"{fpn}spp"] = "neck.spp"
    m[f"{fpn}ghost_5_to_4"] = "neck.fpn4"
    m[f"{fpn}ghost_4_to_3"] = "neck.fpn3"
    m[f"{fpn}stage_3_lane_seg"] = "neck.sa_lane"
    m[f"{fpn}stage_3_semantic_seg"] = "neck.sa_se"
    for n in ("lane", "se"):
        for s in ("3_to_2", "2_to_1", "1_to_0"):
            m[f"{fpn}{n}_seg_ghost_{s}"] = f"neck.{n}_{s}"
    for i in range(8):
        m[f"image_radar_encoder.radar_encoder.rc_blocks.{i}"] = f"radar.block{i}"
    for s in (3, 4, 5):
        m[f"image_radar_encoder.act_stage{s}"] = f"fuse.p{s}"
    return m


def summarize(t, n=512):
    t = t.detach().float().contiguous().flatten()
    step = max(1, t.numel() // n)
    return np.concatenate([[t.mean().item(), t.std().item(), t.abs().max().item(), float(step)],
                           t[::step][:n].numpy()]).astype(np.float32)


def main():
    ns = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() // 2))
    only = set(sys.argv[1:])       # python tests/golden/make_golden.py [config ...]: regenerate just those
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        model = ns.Achelous(phi=cfg["phi"], backbone=cfg["backbone"], neck=cfg.get("neck", "gdf"), **MODEL_KW).eval()
        sd0 = model.state_dict()
        with open(os.path.join(HERE, name + ".keys.json"), "w") as f:
            json.dump({k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd0.items()}, f, indent=0)
        model.load_state_dict(fill_state_dict(sd0, seed=cfg["weight_seed"]), strict=True)
        model.eval()   # achelous.py:171-176 loads the weights and THEN calls .eval(): EfficientFormerV2's Attention4D caches
        #                attention_biases[:, idxs] in .train(False) (ImageEncoder.py:122-128) and would otherwise keep the zeros
        x, xr, pc = make_inputs(2, seed=cfg["input_seed"])
        store = {}
        hooks = []
        tm = tap_map(cfg["phi"], cfg["backbone"])
        mods = dict(model.named_modules())
        for path, tap in tm.items():
            hooks.append(mods[path].register_forward_hook(
                lambda mod, inp, out, tap=tap: store.__setitem__("tap/" + tap, summarize(out))))
        with torch.no_grad():
            det, se, lane, pcs = model(x, xr, pc)
        for h in hooks:
            h.remove()
        for i, d in enumerate(det):
            store[f"det{i}"] = d.numpy()
        store["pc"] = pcs.numpy()
        store["se_argmax"] = se.argmax(1).to(torch.uint8).numpy()
        store["lane_argmax"] = lane.argmax(1).to(torch.uint8).numpy()
        store["se_sub"] = se[:, :, ::4, ::4].contiguous().numpy()
        store["lane_sub"] = lane[:, :, ::4, ::4].contiguous().numpy()
        store["se_sum"] = summarize(se)
        store["lane_sum"] = summarize(lane)

        # ---- post-process goldens: reference decode_outputs (its hard-coded .cuda(rank) made a
        # no-op for this CPU run only) + non_max_suppression (torchvision batched_nms, CPU)
        det_b = [d.clone() for d in det]
        obj_bias = calibrate_obj_bias(det)
        store["obj_bias"] = np.float32(obj_bias)
        for d in det_b:
            d[:, 4] += obj_bias
            d[:, 2:4] += WH_BIAS
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            decoded = ns.decode_outputs([d.clone() for d in det_b], (320, 320), 0)
        finally:
            torch.Tensor.cuda = orig_cuda
        store["decoded"] = decoded.numpy().copy()
        for tag, conf, iou, img_shape, lb in (("a", 0.35, 0.35, (320, 320), False), ("b", 0.25, 0.5, (1080, 1920), True)):
            res = ns.non_max_suppression(decoded.clone(), 7, (320, 320), np.array(img_shape), lb,
                                         conf_thres=conf, nms_thres=iou)
            for b, r in enumerate(res):
                store[f"nms_{tag}_{b}"] = np.zeros((0, 7), np.float32) if r is None else np.asarray(r, np.float32)
                print(name, tag, b, "kept", store[f"nms_{tag}_{b}"].shape[0])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
        print(name, "written", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
