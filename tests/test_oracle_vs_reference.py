"""CPU: the oracle restatement against the UNMODIFIED reference executed live (oracle/ref_loader.py: /root/reference here, the
byte-identical staged copy oracle/_ref where that travelled).  The golden fixtures pin the same thing from stored outputs
(tests/test_oracle_vs_golden.py); this test re-runs the reference itself on fresh seeds, and checks that the fixtures on disk
are still what the reference produces."""
import numpy as np
import pytest
import torch

from achelous_b200.synthetic import make_inputs
from achelous_b200.weights import fill_state_dict
from oracle import functional as OF
from tests.common import GOLDEN_CONFIGS, MODEL_KW, load_golden, neck_of, rel_err

pytestmark = pytest.mark.needs_reference
TOL = 2e-5


def _reference_model(phi, bb, neck, seed):
    from oracle.ref_loader import load_reference
    ns = load_reference()
    model = ns.Achelous(phi=phi, backbone=bb, **dict(MODEL_KW, neck=neck)).eval()
    sd = fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd, strict=True)
    return model.eval(), sd


@pytest.mark.parametrize("phi,bb,neck", [("S0", "en", "gdf"), ("S0", "mv", "gdf"), ("S2", "en", "gdf"), ("S0", "en", "cdf")])
def test_oracle_equals_reference_on_fresh_seeds(phi, bb, neck):
    torch.set_num_threads(4)
    model, sd = _reference_model(phi, bb, neck, seed=41)
    x, xr, pc = make_inputs(1, seed=97)
    with torch.no_grad():
        r_det, r_se, r_lane, r_pc = model(x, xr, pc)
    o_det, o_se, o_lane, o_pc = OF.achelous_forward(sd, x, xr, pc, phi=phi, backbone=bb, neck=neck)
    for a, b in zip(list(o_det) + [o_se, o_lane, o_pc], list(r_det) + [r_se, r_lane, r_pc]):
        assert a.shape == b.shape and rel_err(a, b) < TOL


def test_reference_still_reproduces_the_committed_golden():
    name = "en_gdf_pn_s0"
    phi, bb, wseed, iseed = GOLDEN_CONFIGS[name]
    model, _ = _reference_model(phi, bb, neck_of(name), wseed)
    x, xr, pc = make_inputs(2, seed=iseed)
    with torch.no_grad():
        det, se, lane, pcs = model(x, xr, pc)
    g = load_golden(name)
    for i in range(3):
        assert rel_err(det[i], g[f"det{i}"]) < 1e-6
    assert rel_err(pcs, g["pc"]) < 1e-6
    assert np.array_equal(se.argmax(1).numpy().astype(np.uint8), g["se_argmax"])


def test_reference_state_dict_is_the_contract():
    """the keys.json files the module contract is tested against are the reference's own state_dict()"""
    from tests.common import load_keys
    name = "en_gdf_pn_s0"
    phi, bb, wseed, _ = GOLDEN_CONFIGS[name]
    model, _ = _reference_model(phi, bb, "gdf", wseed)
    spec = load_keys(name)
    sd = model.state_dict()
    assert list(sd) == list(spec)
    assert all(tuple(v.shape) == spec[k][0] and v.dtype == spec[k][1] for k, v in sd.items())
