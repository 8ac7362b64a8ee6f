"""Golden vector for the BUILDER-DEFINED PointNet++ branch (config 4).  The reference contains no PN2 code, so this
fixture is produced by oracle/pn2.py itself and pins the CUDA path (and future oracle edits) to today's definition;
it is NOT evidence of parity with the reference (parity unpinned).

    python tests/golden/make_pn2_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402
from oracle import functional as OF  # noqa: E402
from oracle.pn2 import pointnet2_seg  # noqa: E402

kw = dict(num_det=7, num_seg=9, phi="S2", resolution=320, backbone="en", neck="gdf", pc_seg="pn2", pc_channels=5, pc_classes=8,
          nano_head=True, spp=True)
sd = fill_state_dict(Achelous(**kw).state_dict(), seed=3)
x, xr, pc = make_inputs(3, seed=22)
taps = {}
out = pointnet2_seg(pc[:2], OF.SD(sd, "pc_seg_model."), taps)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pn2_builder_defined.npz"), pc=out.numpy(),
                    fps1=taps["pc.sa1.fps"].numpy(), idx1=taps["pc.sa1.idx"].numpy())
print("written", out.shape)
