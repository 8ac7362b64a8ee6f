"""Device-memory neighbours of one engine tensor (who could overrun into it?)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous
from achelous_b200.synthetic import make_inputs
from achelous_b200.weights import fill_state_dict
kw = dict(num_det=7, num_seg=9, phi="S0", resolution=320, backbone="en", neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8, nano_head=True, spp=True)
model = Achelous(**kw).eval(); model.load_state_dict(fill_state_dict(model.state_dict(), seed=0)); model = model.cuda()
x, xr, pc = [t.cuda() for t in make_inputs(64, seed=1234)]
model(x, xr, pc)
eng = next(iter(model._engines.values()))
items = []
for n, t in eng._bufs.items(): items.append((t.data_ptr(), t.numel() * t.element_size(), "buf " + n))
for i, t in enumerate(eng._keep):
    if isinstance(t, torch.Tensor): items.append((t.data_ptr(), t.numel() * t.element_size(), f"keep[{i}]{tuple(t.shape)}"))
for k, (t, _) in eng._weights.items(): items.append((t.data_ptr(), t.numel() * t.element_size(), "w " + k))
items.sort()
tgt = [i for i, it in enumerate(items) if it[2].startswith(sys.argv[1] if len(sys.argv) > 1 else "keep[24]")][0]
for p, nb, n in items[max(0, tgt - 8): tgt + 4]:
    print(f"{p:#x} .. {p + nb:#x} ({nb:9d} B)  {n}   gap to target start: {items[tgt][0] - (p + nb)}")
