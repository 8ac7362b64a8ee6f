"""Runs selected launches of the forward plan inside a cudaProfilerStart/Stop window so that
`ncu --profile-from-start off --set full ...` captures exactly those kernels (with real plan inputs).

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof \
        python tools/profile_ops.py --ops bb.s0.0.pw1,rc0.deform
"""
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
    ap.add_argument("--ops", required=True, help="comma separated plan op names (see tools/op_times.py output)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--phi", default="S0")
    ap.add_argument("--backbone", default="en")
    ap.add_argument("--neck", default="gdf")
    ap.add_argument("--pc-seg", default="pn")
    ap.add_argument("--rc-simt", action="store_true", help="RCBlocks on the SIMT kernel (A/B against the tensor-core one)")
    a = ap.parse_args()
    kw = dict(num_det=7, num_seg=9, phi=a.phi, resolution=320, backbone=a.backbone, neck=a.neck, pc_seg=a.pc_seg, pc_channels=5,
              pc_classes=8, nano_head=True, spp=True)
    model = Achelous(**kw).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=0))
    model.use_cuda_graph = False
    model.rc_tensor_cores = not a.rc_simt
    model = model.cuda()
    x, xr, pc = [t.cuda() for t in make_inputs(a.batch, seed=1)]
    model(x, xr, pc)
    model(x, xr, pc)
    torch.cuda.synchronize()
    eng = next(iter(model._engines.values()))
    stream = torch.cuda.current_stream().cuda_stream
    want = a.ops.split(",")
    idx = [eng.op_names.index(n) for n in want]
    torch.cuda.cudart().cudaProfilerStart()
    for i in idx:
        fn, args = eng.ops[i]
        _lib.check(fn(*args, stream), eng.op_names[i])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled:", want)


if __name__ == "__main__":
    main()
