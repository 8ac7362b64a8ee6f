"""Which launch stops being reproducible when OTHER kernels share the GPU with it?  Runs a victim op sequence of the B = 64 plan over and
over on one stream while a noise op sequence loops on a second stream, and compares exact checksums of every plan buffer.

    python tools/op_race_probe.py --victim pn.fstn.fill,pn.fstn.c3max --noise rc0.deform [--iters 300]"""
import argparse
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
    ap.add_argument("--victim", required=True)
    ap.add_argument("--noise", default="")
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--noise-reps", type=int, default=2)
    ap.add_argument("--copy-noise", type=int, default=0, help="N x 300 MB device-to-device copies on the noise stream per iteration")
    ap.add_argument("--detail", default="", help="buffer whose differing elements are described")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--backbone", default="en")
    ap.add_argument("--phi", default="S0")
    ap.add_argument("--neck", default="gdf")
    ap.add_argument("--pc-seg", default="pn")
    a = ap.parse_args()
    kw = dict(num_det=7, num_seg=9, phi=a.phi, resolution=320, backbone=a.backbone, neck=a.neck, pc_seg=a.pc_seg, pc_channels=5, pc_classes=8,
              nano_head=True, spp=True)
    model = Achelous(**kw).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=0))
    model.use_cuda_graph = False
    model = model.cuda()
    x, xr, pc = [t.cuda() for t in make_inputs(a.batch, seed=1234)]
    model(x, xr, pc)
    eng = next(iter(model._engines.values()))
    eng.multi_stream = False
    eng.forward_static()          # every buffer holds its single-stream value
    torch.cuda.synchronize()
    names = [n for n in eng._bufs if not n.startswith("in.")]
    bufs = [eng._bufs[n] for n in names]
    for i, t in enumerate(eng._keep):
        if isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == a.batch and t.numel() < (1 << 22):
            names.append(f"keep[{i}]{tuple(t.shape)}")
            bufs.append(t)

    def sums():
        return torch.stack([b.contiguous().view(torch.int32).sum(dtype=torch.int64) for b in bufs])

    def run(ops, stream):
        for n in ops:
            i = eng.op_names.index(n)
            fn, args = eng.ops[i]
            _lib.check(fn(*args, stream.cuda_stream), n)

    def expand(spec):      # names, or prefixes ending in '*'
        out = []
        for tok in [t for t in spec.split(",") if t]:
            out += [n for n in eng.op_names if n.startswith(tok[:-1])] if tok.endswith("*") else [tok]
        return out
    victim = expand(a.victim)
    noise = expand(a.noise)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(sa):
        run(victim, sa)
        ref = sums()
    torch.cuda.synchronize()
    bad = {}
    na = torch.empty(75_000_000, device="cuda") if a.copy_noise else None
    nb = torch.empty_like(na) if a.copy_noise else None
    det = bufs[names.index(a.detail)] if a.detail else None
    det_ref = det.clone() if det is not None else None
    shown = 0
    for it in range(a.iters):
        if a.copy_noise:
            with torch.cuda.stream(sb):
                for _ in range(a.copy_noise):
                    nb.copy_(na, non_blocking=True)
        if noise:
            with torch.cuda.stream(sb):
                for _ in range(a.noise_reps):
                    run(noise, sb)
        with torch.cuda.stream(sa):
            run(victim, sa)
            s = sums()
        torch.cuda.synchronize()
        for i in (s != ref).nonzero().flatten().tolist():
            bad.setdefault(names[i], []).append(it)
        if det is not None and shown < 5 and not torch.equal(det, det_ref):
            shown += 1
            ne = (det != det_ref)
            idx = ne.nonzero()
            W_ = det.shape[3]
            pix = (idx[:, 2] * W_ + idx[:, 3])
            lanes = sorted(set((pix % 128).tolist()))
            runs, st = [], None
            for v in lanes + [None]:
                if st is None:
                    st = prev = v
                elif v is not None and v == prev + 1:
                    prev = v
                else:
                    runs.append(f"{st}-{prev}")
                    st = prev = v
            print(f"  iter {it}: {int(ne.sum())} elements differ; frames {sorted(set(idx[:, 0].tolist()))[:10]}, channels {int(idx[:, 1].min())}..{int(idx[:, 1].max())} "
                  f"({len(set(idx[:, 1].tolist()))} distinct), pixel tiles {sorted(set((pix // 128).tolist()))}, pixel-in-tile runs {runs}; "
                  f"max |diff| {(det - det_ref)[ne].abs().max().item():.3e}; ref {det_ref[ne][:3].tolist()} cur {det[ne][:3].tolist()}")
    tag = f"victim {a.victim} | noise {a.noise}"
    if not bad:
        print(f"OK   {tag}: {a.iters} runs identical")
    else:
        print(f"RACE {tag}: " + ", ".join(f"{n} x{len(v)}" for n, v in bad.items()))


if __name__ == "__main__":
    main()
