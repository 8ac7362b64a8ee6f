"""CPU, world_size 2, gloo: the sharded-frames + single all-gather scheme of bench.py --gpus N
(rank r owns frames [r*B, (r+1)*B); packed per-frame output rows are all-gathered once)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, B, F, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234 + rank)
    local = torch.randn(B, F, generator=g)           # stands for Engine.packed_out of this rank
    gathered = torch.empty(world * B, F)
    dist.all_gather_into_tensor(gathered, local)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)         # max-over-ranks timing reduction used by bench.py
    if rank == 0:
        ret["gathered"] = gathered.clone()
        ret["tmax"] = t.item()
    dist.destroy_process_group()


def test_frame_sharding_allgather_equals_single_process():
    world, B, F = 2, 3, 37
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29531, B, F, ret), nprocs=world, join=True)
    expect = torch.cat([torch.randn(B, F, generator=torch.Generator().manual_seed(1234 + r)) for r in range(world)])
    assert torch.equal(ret["gathered"], expect)
    assert ret["tmax"] == float(world)
