"""CPU, world_size 2, gloo: the sharded-frames + single in-place all-gather scheme of bench.py --gpus N.
Each rank runs the REAL engine plan (host logic: weight folding, views, output packing; kernels through the C-ABI emulator) on
its own frames, writing straight into its rows of the gather buffer; after ONE all-gather every rank holds all frames' packed
outputs, and they equal a single-process run of the same frames bit for bit - raw fp32 outputs and the compact record."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import MODEL_KW

B = 2
SPEC = dict(conf_thres=0.05, nms_thres=0.5, max_det=32)


def _engine(model, compact, out=None):
    from achelous_b200.engine import Engine
    return Engine(model, B, "cpu", dry_run=True, compact=SPEC if compact else None, out=out)


def _run(eng, seed):
    from achelous_b200.synthetic import make_inputs
    from tests.abi_emulator import emulate_engine
    for dst, src in zip(eng.input_tensors(), make_inputs(B, seed=seed)):
        dst.copy_(src)
    emulate_engine(eng)
    return eng.packed_out


def _model():
    from achelous_b200.nets.Achelous import Achelous
    from achelous_b200.weights import fill_state_dict
    torch.set_num_threads(2)
    m = Achelous(phi="S0", backbone="en", **MODEL_KW).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=2))
    return m


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    for compact in (False, True):
        probe = _engine(model, compact)
        gathered = torch.zeros(world * B, probe.packed_out.shape[1], dtype=probe.packed_out.dtype)
        eng = _engine(model, compact, out=gathered[rank * B:(rank + 1) * B])     # the plan writes into this rank's rows
        assert eng.packed_out.data_ptr() == gathered[rank * B].data_ptr()
        _run(eng, 1234 + rank)
        dist.all_gather_into_tensor(gathered, eng.packed_out.clone())           # gloo wants a non-aliased input; NCCL gathers in place
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                                 # max-over-ranks timing reduction used by bench.py
        if rank == 0:
            ret["compact" if compact else "raw"] = gathered.clone()
            ret["tmax"] = t.item()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_sharded_engines_allgather_equals_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29531, ret), nprocs=world, join=True)
    assert ret["tmax"] == float(world)
    model = _model()
    for compact in (False, True):
        eng = _engine(model, compact)
        expect = torch.cat([_run(eng, 1234 + r).clone() for r in range(world)])
        got = ret["compact" if compact else "raw"]
        assert got.dtype == expect.dtype and torch.equal(got, expect)
        views = eng.unpack(got)                                                   # gathered rows unpack like a local buffer
        if compact:
            assert views.se_mask.shape == (world * B, 320, 320) and int(views.det_count.min()) >= 0
        else:
            assert views[1].shape == (world * B, 9, 320, 320)
