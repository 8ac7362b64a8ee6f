"""Where do the wrong entries of PointNet's global max feature come from?  On a replay whose fstn global feature differs from the first
replay, every wrong (frame, output) value is searched among relu(scale * W a + bias) of ALL frames / points."""
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
g = [t for t in eng._keep if isinstance(t, torch.Tensor) and t.dtype == torch.float32 and tuple(t.shape) == (64, 1024)][1]   # fstn global feature
a2 = eng._bufs["pn.fstn.c2"].view(64, 128, 512)
W = eng._weights["pn.fstn.c3.wt"][0][:, :1024]          # (128, 1024) K-major
sc, bi = eng._weights["pn.fstn.c3.s"][0], eng._weights["pn.fstn.c3.b"][0]
eng.forward_static(); torch.cuda.synchronize()
ref = g.clone()
Y = torch.relu(torch.einsum("ko,bkp->bop", W.double(), a2.double()) * sc.double()[None, :, None] + bi.double()[None, :, None]).float()   # (64, 1024, 512)
print("reference vs recomputed max: max abs diff", (Y.amax(2) - ref).abs().max().item())
shown = 0
for it in range(400):
    eng.forward_static(); torch.cuda.synchronize()
    ne = (g != ref)
    if not ne.any():
        continue
    idx = ne.nonzero()
    print(f"replay {it}: {len(idx)} wrong entries; frames {sorted(set(idx[:,0].tolist()))}")
    for b, o in idx[:10].tolist():
        v = g[b, o].item()
        hit = ((Y[:, o, :] - v).abs() < 2e-6 * max(1.0, abs(v))).nonzero()
        hit_any_o = ((Y[b] - v).abs() < 2e-6 * max(1.0, abs(v))).nonzero()
        print(f"   g[{b},{o}] = {v:.6f} (truth {ref[b,o].item():.6f}); same output o, (frame, point) with that value: {hit[:6].tolist()}; "
              f"same frame, (output, point): {hit_any_o[:6].tolist()}")
    shown += 1
    if shown >= 3:
        break
