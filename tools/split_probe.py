"""Does running the batch as n concurrent sub-batches (independent graph replays on n streams) beat one big batch?"""
import sys, os, copy, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous
from achelous_b200.synthetic import make_inputs
from achelous_b200.weights import fill_state_dict

kw = dict(num_det=7, num_seg=9, phi="S0", resolution=320, backbone="en", neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8, nano_head=True, spp=True)
B = 64
def mk():
    m = Achelous(**kw).eval(); m.load_state_dict(fill_state_dict(m.state_dict(), seed=0)); return m.cuda()
for n in (1, 2, 4):
    models = [mk() for _ in range(n)]
    x, xr, pc = [t.cuda() for t in make_inputs(B // n, seed=1)]
    engs = []
    for m in models:
        m(x, xr, pc)
        engs.append(next(iter(m._engines.values())))
    streams = [torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize()
    def step():
        main = torch.cuda.current_stream()
        for s, e in zip(streams, engs):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                e.graph.replay()
        for s in streams:
            main.wait_stream(s)
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"n_split={n}: {ms:.3f} ms per {B} frames -> {B/ms*1e3:.0f} frames/s")
