"""Software pipeline across steps: the batch as n sub-batches, each an engine replaying its CUDA graph on its OWN stream without a
per-step join, the streams started `offset` apart - does the low-resolution tail of one sub-batch (backbone stages 2-3, SPP, FPN:
small grids, nothing else left to overlap with) hide under the high-resolution head of the other?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402

kw = dict(num_det=7, num_seg=9, phi="S0", resolution=320, backbone="en", neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8,
          nano_head=True, spp=True)
B, STEPS = 64, 40


def mk():
    m = Achelous(**kw).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    return m.cuda()


clock_khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000
for n, offset_ms in ((1, 0.0), (2, 0.0), (2, 1.0), (2, 2.0), (2, 3.0), (4, 0.0), (4, 1.0)):
    models = [mk() for _ in range(n)]
    x, xr, pc = [t.cuda() for t in make_inputs(B // n, seed=1)]
    engs = []
    for m in models:
        m(x, xr, pc)
        engs.append(next(iter(m._engines.values())))
    streams = [torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize()

    def run(steps):
        main = torch.cuda.current_stream()
        for i, (s, e) in enumerate(zip(streams, engs)):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                if i and offset_ms:
                    torch.cuda._sleep(int(i * offset_ms * 1e-3 * 1.965e9))
                for _ in range(steps):
                    e.graph.replay()
        for s in streams:
            main.wait_stream(s)

    run(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(STEPS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    print(f"n_split={n} offset={offset_ms} ms: {ms:.3f} ms per {B} frames -> {B / ms * 1e3:.0f} frames/s", flush=True)
    del models, engs
    torch.cuda.empty_cache()
