"""Where the end-to-end step goes: copy bandwidths alone and concurrently, and the forward alone vs under DMA load."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402

kw = dict(num_det=7, num_seg=9, phi="S0", resolution=320, backbone="en", neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8,
          nano_head=True, spp=True)
model = Achelous(**kw).eval()
model.load_state_dict(fill_state_dict(model.state_dict(), seed=0))
model = model.cuda()
B = 64
x, xr, pc = [t.cuda() for t in make_inputs(B, seed=1)]
model(x, xr, pc)
eng = next(iter(model._engines.values()))
hin = torch.empty(158 * 1024 * 1024 // 4).pin_memory()
din = torch.empty_like(hin, device="cuda")
hout = torch.empty(B, eng.frame_elems).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(eng.packed_out, non_blocking=True)


print("forward alone      %.2f ms" % timeit(eng.forward_static))
print("H2D 158 MB alone   %.2f ms" % timeit(h2d))
print("D2H 296 MB alone   %.2f ms" % timeit(d2h))
print("H2D + D2H together %.2f ms" % timeit(lambda: (h2d(), d2h())))
print("forward + D2H      %.2f ms" % timeit(lambda: (eng.forward_static(), d2h())))
print("forward + H2D      %.2f ms" % timeit(lambda: (eng.forward_static(), h2d())))
print("forward + both     %.2f ms" % timeit(lambda: (eng.forward_static(), h2d(), d2h())))
