"""PeerGather on N GPUs (torchrun): correctness against NCCL's all-gather over many steps, then the time of one gather of
`--mb` megabytes per rank with nothing else running (copy engines over NVLink) next to NCCL's.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_probe.py"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.peer_gather import PeerGather  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=296.0)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--copy-streams", type=int, default=2)
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)

    # ---- correctness: 9 steps through 2 slots, ragged widths, fp32 and uint8 rows, consumer = a checksum kernel before release
    for dtype, rows, width in ((torch.float32, 5, 1031), (torch.uint8, 3, 212496)):
        pg = PeerGather(rows, width, dtype, dev, copy_streams=a.copy_streams)
        comm = torch.cuda.Stream(dev)
        ok = True
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for i in range(9):
            s, seq = i % 2, i + 1
            mine = (torch.rand(rows, width, device=dev, generator=g) * 200).to(dtype)
            pg.slot(s)[pg.my_rows].copy_(mine)
            ref = torch.empty(world * rows, width, device=dev, dtype=dtype)
            dist.all_gather_into_tensor(ref, mine)
            comm.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(comm):
                pg.push(s, seq, comm)
                pg.wait(s, seq, comm)
                got = pg.slot(s).clone()
                pg.release(s, seq, comm)
            torch.cuda.current_stream(dev).wait_stream(comm)
            ok &= bool(torch.equal(got, ref))
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        assert t.item() == 1, f"PeerGather mismatch ({dtype})"
        pg.close()

    # ---- time of one gather alone
    width = 1 << 18
    rows = max(1, int(a.mb * (1 << 20) / 4 / width))
    pg = PeerGather(rows, width, torch.float32, dev, copy_streams=a.copy_streams)
    nccl_out = torch.empty(world * rows, width, device=dev)
    mine = pg.slot(0)[pg.my_rows]
    comm = torch.cuda.current_stream(dev)
    res = {}
    for name in ("peer", "nccl"):
        seq = 0
        times = []
        for it in range(a.steps + 3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "peer":
                seq += 1
                s = (seq - 1) % 2
                pg.push(s, seq, comm)
                pg.wait(s, seq, comm)
                pg.release(s, seq, comm)
            else:
                dist.all_gather_into_tensor(nccl_out, mine)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(times)[len(times) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = t.item()
    if rank == 0:
        part = rows * width * 4
        print(json.dumps({"world": world, "mb_per_rank": part / 2**20, "copy_streams": a.copy_streams,
                          "peer_ms": res["peer"], "nccl_ms": res["nccl"],
                          "peer_in_GBps_per_rank": (world - 1) * part / res["peer"] / 1e6,
                          "nccl_in_GBps_per_rank": (world - 1) * part / res["nccl"] / 1e6, "check": "ok"}), flush=True)
    pg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
