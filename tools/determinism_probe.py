"""Replays the B = 64 plan many times on the same inputs and compares an exact integer checksum of EVERY plan buffer with the
first replay: a run-to-run difference names the first launch (in plan order) whose output is not reproducible.

    python tools/determinism_probe.py [--iters 300] [--copy-noise]   (--copy-noise: a 300 MB device-to-device copy loop on a side
                                                                      stream, to perturb timing like the multi-GPU gather does)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--backbone", default="en")
    ap.add_argument("--phi", default="S0")
    ap.add_argument("--neck", default="gdf")
    ap.add_argument("--pc-seg", default="pn")
    ap.add_argument("--copy-noise", action="store_true")
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--merge-lanes", default="", help="e.g. '2>0,3>0': ops of lane 2 / 3 are issued on lane 0 instead (bisecting which overlap matters)")
    ap.add_argument("--only", default="", help="comma separated op-name prefixes: launch ONLY these ops (e.g. 'pn.,rc0.') - inputs of the others keep "
                    "the values of the first full run")
    ap.add_argument("--no-conv3-tc", action="store_true", help="dense 3x3 convs on conv_dense (CUDA cores) instead of conv3x3_tc")
    ap.add_argument("--detail", default="", help="buffer name (as printed) whose differing elements are described")
    ap.add_argument("--single-stream", action="store_true", help="eager launches on ONE stream (no lanes): separates missing cross-lane "
                    "dependencies from races inside a kernel")
    a = ap.parse_args()
    kw = dict(num_det=7, num_seg=9, phi=a.phi, resolution=320, backbone=a.backbone, neck=a.neck, pc_seg=a.pc_seg, pc_channels=5, pc_classes=8,
              nano_head=True, spp=True)
    model = Achelous(**kw).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=0))
    model.use_cuda_graph = not (a.eager or a.single_stream)
    model.conv3_tensor_cores = not a.no_conv3_tc
    model = model.cuda()
    x, xr, pc = [t.cuda() for t in make_inputs(a.batch, seed=1234)]
    model(x, xr, pc)
    eng = next(iter(model._engines.values()))
    if a.single_stream:
        eng.multi_stream = False
    if a.merge_lanes:
        mp = {int(x.split(">")[0]): int(x.split(">")[1]) for x in a.merge_lanes.split(",")}
        eng.op_lane = [mp.get(l, l) for l in eng.op_lane]
        eng.sync_before = {i: [(mp.get(w, w), mp.get(v, v)) for w, v in lst if mp.get(w, w) != mp.get(v, v)] for i, lst in eng.sync_before.items()}
        eng.sync_end = [(mp.get(w, w), mp.get(v, v)) for w, v in eng.sync_end if mp.get(w, w) != mp.get(v, v)]
        eng.graph = None
    if a.only:
        pre = tuple(a.only.split(","))
        keep = [i for i, n in enumerate(eng.op_names) if n.startswith(pre)]
        remap = {old: new for new, old in enumerate(keep)}
        sb = {}
        for i, lst in eng.sync_before.items():          # a dependency recorded before a dropped op moves to the next kept op
            nxt = [k for k in keep if k >= i]
            if nxt:
                sb.setdefault(remap[nxt[0]], []).extend(lst)
        eng.ops = [eng.ops[i] for i in keep]
        eng.op_names = [eng.op_names[i] for i in keep]
        eng.op_lane = [eng.op_lane[i] for i in keep]
        lanes = set(eng.op_lane) | {0}
        eng.sync_before = {i: [(w, v) for w, v in lst if w in lanes and v in lanes] for i, lst in sb.items()}
        eng.sync_end = [(w, v) for w, v in eng.sync_end if w in lanes and v in lanes]
        eng.graph = None
        print("launching only:", len(eng.ops), "ops")
    names = [n for n in eng._bufs if not n.startswith("in.")]
    bufs = [eng._bufs[n] for n in names]
    # small per-frame vectors (global max features, FC outputs, ECA means) live in eng._keep without a name
    for i, t in enumerate(eng._keep):
        if isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == a.batch and t.numel() < (1 << 22):
            names.append(f"keep[{i}]{tuple(t.shape)}")
            bufs.append(t)
    bufs.append(eng.packed_out)
    names.append("packed_out")

    def sums():
        return torch.stack([b.contiguous().view(torch.int32).sum(dtype=torch.int64) for b in bufs])

    side = torch.cuda.Stream()
    noise_a = torch.empty(75_000_000, device="cuda") if a.copy_noise else None
    noise_b = torch.empty_like(noise_a) if a.copy_noise else None
    eng.forward_static()
    torch.cuda.synchronize()
    ref = sums()
    bad = {}
    det_i = names.index(a.detail) if a.detail else None
    det_ref = bufs[det_i].clone() if det_i is not None else None
    shown = 0
    for it in range(a.iters):
        if a.copy_noise:
            with torch.cuda.stream(side):
                for _ in range(3):
                    noise_b.copy_(noise_a, non_blocking=True)
        eng.forward_static()
        s = sums()
        torch.cuda.synchronize()
        diff = (s != ref).nonzero().flatten().tolist()
        if det_i is not None and det_i in diff and shown < 6:
            shown += 1
            cur = bufs[det_i]
            ne = (cur != det_ref)
            idx = ne.nonzero()
            d = (cur - det_ref)[ne]
            rows = sorted(set(idx[:, 0].tolist()))
            cols = idx[:, 1].tolist() if idx.shape[1] > 1 else []
            print(f"  replay {it}: {int(ne.sum())} elements differ, rows {rows[:12]}, cols {min(cols) if cols else None}..{max(cols) if cols else None}, "
                  f"cur-ref min {d.min().item():.3e} max {d.max().item():.3e}; ref values there {det_ref[ne][:4].tolist()} cur {cur[ne][:4].tolist()}")
        for i in diff:
            bad.setdefault(names[i], []).append(it)
    if not bad:
        print(f"deterministic: {a.iters} replays, {len(bufs)} buffers identical")
        return
    order = {n: i for i, n in enumerate(names)}
    print("keep tensors, in creation order:", [n for n in names if n.startswith("keep[")])
    print(f"NON-DETERMINISTIC buffers ({a.iters} replays):")
    for n in sorted(bad, key=lambda n: order[n]):
        print(f"  {n:40s} differs in {len(bad[n])} replays (first at {bad[n][0]})")


if __name__ == "__main__":
    main()
