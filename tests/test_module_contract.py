"""CPU: the drop-in boundary of nets/Achelous.py - state-dict contract, wrappers the module must survive
(deepcopy, children()/attribute pokes, pickling, nn.DataParallel replicas), stale-weight detection, constructor validation."""
import copy
import io
import pickle

import pytest
import torch
import torch.nn as nn

from achelous_b200.engine import Engine
from achelous_b200.nets.Achelous import Achelous, Achelous3T
from achelous_b200.weights import fill_state_dict
from tests.common import GOLDEN_CONFIGS, MODEL_KW, load_keys, neck_of


def _model(phi="S0", bb="en", neck="gdf", seed=1):
    m = Achelous(phi=phi, backbone=bb, **dict(MODEL_KW, neck=neck)).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=seed))
    return m


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_state_dict_contract(name):
    """keys, ORDER, shapes and dtypes equal the reference's state_dict() (tests/golden/*.keys.json, written by
    tests/golden/make_golden.py from the unmodified reference): achelous.py:171 does a strict load of reference checkpoints"""
    phi, bb, _, _ = GOLDEN_CONFIGS[name]
    spec = load_keys(name)
    sd = Achelous(phi=phi, backbone=bb, **dict(MODEL_KW, neck=neck_of(name))).state_dict()
    assert list(sd.keys()) == list(spec.keys())
    for k, (shape, dtype) in spec.items():
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dtype, k


def test_deepcopy_children_and_pickle():
    m = _model()
    m._engines[("fake",)] = object()
    e = copy.deepcopy(m)                        # ModelEMA (detection_loss.py:441)
    assert e._engines == {} and e._owner() is e and m._owner() is m
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(m.state_dict().values(), e.state_dict().values()))
    for child in m.modules():                   # utils/callbacks.py:151-159 pokes attributes on every module
        child.deploy = True
    assert len(list(m.children())) == 3
    m._engines.clear()
    m2 = pickle.loads(pickle.dumps(m))          # torch.save(model) style whole-module pickling
    assert m2._owner() is m2 and list(m2.state_dict()) == list(m.state_dict())
    buf = io.BytesIO()
    torch.save(m.state_dict(), buf)
    buf.seek(0)
    m2.load_state_dict(torch.load(buf))


def _replicate_on_cpu(module):
    """torch/nn/parallel/replicate.py without the device broadcast: shallow module copies with EMPTY `_parameters`
    (former parameters become plain tensor attributes)"""
    mods = list(module.modules())
    idx = {m: i for i, m in enumerate(mods)}
    reps = [m._replicate_for_data_parallel() for m in mods]
    for m, r in zip(mods, reps):
        for key, child in m._modules.items():
            r._modules[key] = None if child is None else reps[idx[child]]
        for key, p in m._parameters.items():
            if p is not None:
                setattr(r, key, p.detach().clone())
        for key, b in m._buffers.items():
            r._buffers[key] = None if b is None else b.clone()
    return reps[0]


def test_dataparallel_replica_resolves_its_owner():
    """nn.DataParallel (achelous.py:176) calls forward on replicas that own no parameters: they must find the wrapped module,
    share its engine cache and build plans from ITS parameters"""
    m = _model()
    r = _replicate_on_cpu(m)
    assert next(r.parameters(), None) is None           # what broke round 1: next(self.parameters())
    assert r._owner() is m and r._engines is m._engines
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(torch.zeros(1, 3, 320, 320), torch.zeros(1, 3, 320, 320), torch.zeros(1, 5, 512))
    eng = Engine(r._owner(), 1, "cpu", dry_run=True, n_points=512)
    assert len(eng._params) == len(m.state_dict())
    wrapped = nn.DataParallel(m, device_ids=None) if torch.cuda.is_available() else None
    assert wrapped is None or wrapped.module is m


def test_stale_weight_detection():
    m = _model()
    eng = Engine(m, 1, "cpu", dry_run=True)
    key = "det0.stem.wt"
    w0 = eng._weights[key][0].clone()
    assert not eng.refresh_weights()
    # (1) in-place op on the parameter: autograd version counter
    with torch.no_grad():
        m.det_head.stems[0].conv.weight.mul_(2.0)
    assert eng.refresh_weights() and torch.allclose(eng._weights[key][0], 2 * w0)
    # (2) load_state_dict(assign=True) swaps the Parameter objects: caught by the post-hook's epoch, looked up by name
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["det_head.stems.0.conv.weight"] = sd["det_head.stems.0.conv.weight"] * 0.25
    m.load_state_dict(sd, assign=True)
    assert eng.refresh_weights() and torch.allclose(eng._weights[key][0], 0.5 * w0)
    # (3) a write through .data is invisible to both: documented, needs invalidate()
    m.det_head.stems[0].conv.weight.data.mul_(2.0)
    assert not eng.refresh_weights()
    m.invalidate()
    assert eng.refresh_weights() and torch.allclose(eng._weights[key][0], w0)
    # (4) replacing a parameter object by assignment
    m.det_head.stems[0].conv.weight = nn.Parameter(m.det_head.stems[0].conv.weight.detach() * 3.0)
    m.invalidate()
    assert eng.refresh_weights() and torch.allclose(eng._weights[key][0], 3 * w0)


def test_constructor_validation():
    with pytest.raises(ValueError, match="multiple of 64"):
        Achelous(phi="S0", backbone="en", **dict(MODEL_KW, resolution=416))     # the reference's default (13 x 13 at stride 32)
    with pytest.raises(ValueError, match="multiple of 64"):
        Achelous3T(7, 9, resolution=352)
    Achelous(phi="S0", backbone="en", **dict(MODEL_KW, resolution=384))
    with pytest.raises(NotImplementedError):
        Achelous(phi="S0", backbone="en", **MODEL_KW).train()(torch.zeros(1, 3, 320, 320), torch.zeros(1, 3, 320, 320), torch.zeros(1, 5, 512))


def test_compact_options_validated():
    m = _model()
    with pytest.raises(ValueError):
        m._compact_key("packed", {})
    with pytest.raises(TypeError):
        m._compact_key("raw", {"max_det": 3})
    with pytest.raises(TypeError):
        m._compact_key("compact", {"bogus": 1})
    assert m._compact_key("compact", {"keep_classes": [8, 0]}) == m._compact_key("compact", {"keep_classes": (0, 8)})
