"""Per-launch device times of the forward plan (CUDA events around every C-ABI call, eager mode).
    python tools/op_times.py [--batch 64] [--phi S0] [--backbone en] [--out gpurun_out/op_times.json]
Used to rank kernels for optimisation; bench.py is the number of record."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200 import _lib  # noqa: E402
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--phi", default="S0")
    ap.add_argument("--backbone", default="en")
    ap.add_argument("--neck", default="gdf")
    ap.add_argument("--pc-seg", default="pn")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tc", default="default", help="default | all | off")
    ap.add_argument("--compact", action="store_true", help="time the outputs='compact' plan")
    ap.add_argument("--mlp-group-mb", type=float, default=None, help="override model.mlp_group_bytes (0: whole batch)")
    ap.add_argument("--out", default="gpurun_out/op_times.json")
    a = ap.parse_args()
    kw = dict(num_det=7, num_seg=9, phi=a.phi, resolution=320, backbone=a.backbone, neck=a.neck, pc_seg=a.pc_seg, pc_channels=5,
              pc_classes=8, nano_head=True, spp=True)
    model = Achelous(**kw).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=0))
    model.use_cuda_graph = False
    if a.mlp_group_mb is not None:
        model.mlp_group_bytes = int(a.mlp_group_mb * (1 << 20))
    if a.tc != "default":
        model.use_tensor_cores = "all" if a.tc == "all" else False
    model = model.cuda()
    x, xr, pc = [t.cuda() for t in make_inputs(a.batch, seed=1)]
    model(x, xr, pc, outputs="compact" if a.compact else "raw")
    torch.cuda.synchronize()
    eng = next(iter(model._engines.values()))
    stream = torch.cuda.current_stream().cuda_stream
    n = len(eng.ops)
    times = [[] for _ in range(n)]
    for _ in range(a.reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i, (fn, args) in enumerate(eng.ops):
            _lib.check(fn(*args, stream), eng.op_names[i])
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(n):
            times[i].append(evs[i].elapsed_time(evs[i + 1]))
    med = [sorted(t)[len(t) // 2] for t in times]
    total = sum(med)
    rows = sorted(zip(eng.op_names, [fn.__name__ for fn, _ in eng.ops], med), key=lambda r: -r[2])
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"batch": a.batch, "total_ms": total, "ops": [dict(name=r[0], fn=r[1], ms=r[2]) for r in rows]}, f, indent=1)
    print(f"B={a.batch} sum of per-op medians {total:.3f} ms -> {a.batch / total * 1e3:.0f} frames/s (eager, event-serialised)")
    by_fn = {}
    for r in rows:
        by_fn[r[1]] = by_fn.get(r[1], 0.0) + r[2]
    for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {v:8.3f} ms  {100 * v / total:5.1f}%")
    for r in rows[:25]:
        print(f"  {r[0]:34s} {r[1]:22s} {r[2]:8.3f} ms")
    # whole-plan time, graph replay
    model.use_cuda_graph = True
    eng.use_graph = True
    for _ in range(3):
        eng.forward_static()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward_static()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"graph replay: {ms:.3f} ms / step -> {a.batch / ms * 1e3:.0f} frames/s")


if __name__ == "__main__":
    main()
