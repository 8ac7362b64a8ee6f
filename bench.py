#!/usr/bin/env python
"""Headline benchmark: frames/sec of the full 5-task Achelous forward (320x320 RGB + 320x320 radar map + 512 points) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config en_s0|mv_s0|en_s2_pn2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).
  value        device-timed throughput of the RAW forward (the reference's return value: fp32 logits, the parity artefact), inputs
               resident in HBM, CUDA-graph replay of the whole launch plan [+ one in-place all-gather of the packed outputs (copy-engine pushes over NVLink peer memory)
               when N > 1].  K steps are timed `--repeats` times (barrier + synchronize on both sides, max over ranks each time);
               the MEDIAN repeat is reported.
  compact      the same loop with outputs="compact" (decode + NMS rows, uint8 class maps, point classes produced inside the
               plan - what achelous.py:259-297 derives from the logits; 0.23 MB instead of 4.6 MB per frame to gather)
  e2e          the same metric through the public nn.Module surface with pinned HOST inputs and a device->host read of the
               results inside the timed region: Achelous.stream_forward(compact=True); `raw_logits_value` = the same with the
               raw fp32 outputs copied out (PCIe-bound)
  roofline     the dominant kernel timed live with CUDA events
  cpu_baseline the reference's own PyTorch forward (oracle/_ref = staged unmodified reference files; the oracle port when that
               is absent or, for config 4, because the reference ships no PointNet++) on the box's host cores
`--impl reference` times that CPU implementation alone, same metric / config."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec full 5-task forward @320x320+512pts"
UNIT = "frames/s"
BASE_KW = dict(num_det=7, num_seg=9, resolution=320, neck="gdf", pc_channels=5, pc_classes=8, nano_head=True, spp=True)
CONFIGS = {   # BASELINE.json configs[1], [2], [3]
    "en_s0": dict(kw=dict(BASE_KW, phi="S0", backbone="en", pc_seg="pn"), batch=64,
                  workload="EN-GDF-PN-S0 inference, 320x320 RGB + radar map + 512 pts"),
    "mv_s0": dict(kw=dict(BASE_KW, phi="S0", backbone="mv", pc_seg="pn"), batch=64,
                  workload="MV-GDF-PN-S0 inference (MobileViT backbone), 320x320 RGB + radar map + 512 pts"),
    "en_s2_pn2": dict(kw=dict(BASE_KW, phi="S2", backbone="en", pc_seg="pn2"), batch=32,
                      workload="EN-GDF-PN2-S2 inference (PointNet++ head: builder-defined, the reference ships none), 320x320 RGB + radar map + 512 pts"),
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arm
class CpuForward:
    """The reference's CPU forward for one config: the UNMODIFIED reference module (kind "reference": /root/reference or the staged
    copy oracle/_ref) when it can be imported and implements the config, else the oracle port (kind "port")."""

    def __init__(self, cfg_name):
        import torch
        from achelous_b200.nets.Achelous import Achelous
        from achelous_b200.weights import fill_state_dict
        cfg = CONFIGS[cfg_name]
        kw = cfg["kw"]
        self.kw = kw
        self.kind, self.how, self.model = "port", "oracle/functional.py (PyTorch CPU fp32 restatement)", None
        spec = Achelous(**kw).state_dict()
        self.sd = fill_state_dict(spec, seed=0)
        if kw["pc_seg"] == "pn":       # the reference has no PointNet++: config 4 can only be timed as the port
            try:
                from oracle.ref_loader import load_reference, reference_available, reference_kind
                if reference_available():
                    import contextlib
                    with contextlib.redirect_stdout(sys.stderr):    # the reference's imports print hints: stdout carries ONE JSON line
                        ns = load_reference()
                        m = ns.Achelous(**kw).eval()
                    m.load_state_dict(self.sd, strict=True)
                    self.model, self.kind = m.eval(), "reference"
                    self.how = f"unmodified reference nets/Achelous.py:49-53 ({reference_kind()}), PyTorch CPU fp32, torch.no_grad()"
            except Exception as e:   # pragma: no cover
                self.how += f" [reference import failed: {type(e).__name__}: {e}]"
        self.torch = torch

    def __call__(self, x, xr, pc):
        if self.model is not None:
            with self.torch.no_grad():
                return self.model(x, xr, pc)
        from oracle import functional as OF
        return OF.achelous_forward(self.sd, x, xr, pc, phi=self.kw["phi"], backbone=self.kw["backbone"], pc_seg=self.kw["pc_seg"])

    def time(self, batch, threads, warmup, steps):
        from achelous_b200.synthetic import make_inputs
        self.torch.set_num_threads(threads)
        x, xr, pc = make_inputs(batch, seed=1234)
        ts = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            self(x, xr, pc)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
        return ts


def thread_counts(cores):
    n, out = 1, []
    while n < cores:
        out.append(n)
        n *= 2
    return out + [cores]


def cpu_sweep(fwd, cores, warm, timed, budget_s):
    """B=1, torch threads swept over 1, 2, 4, ..., cores; best median kept.  Also tries one batched point
    (B=8, all cores) - a throughput-minded CPU user would batch.  Stops adding points when the time budget is spent."""
    t_start, tried, best = time.perf_counter(), [], None
    for batch, th in [(1, n) for n in thread_counts(cores)] + [(8, cores)]:
        if time.perf_counter() - t_start > budget_s and best is not None:
            break
        ts = fwd.time(batch, th, warm, timed)
        fps = batch / statistics.median(ts)
        tried.append({"batch": batch, "threads": th, "frames_per_s": round(fps, 3)})
        if best is None or fps > best[0]:
            best = (fps, batch, th)
    return best, tried


def run_reference(args):
    """CPU arm: the reference's own forward on the host cores, best (batch, threads) point of a short sweep, then K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fwd = CpuForward(args.config)
    best, tried = cpu_sweep(fwd, cores, 1, 2, budget_s=60.0)
    _, batch, th = best
    ts = fwd.time(batch, th, max(args.warmup, 1), args.steps)
    total = sum(ts)
    value = batch * len(ts) / total
    cfg = CONFIGS[args.config]
    sample = (f"{len(ts)} steps x {batch} frame(s) at {th} torch threads (best point of a sweep over B=1 x threads {thread_counts(cores)} and "
              f"B=8 x {cores}); {fwd.how}")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": cfg["workload"], "config": args.config, "batch_per_step": batch, "device": "host CPU", "torch_threads": th,
                       "host_cores": cores, "sweep": tried},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": fwd.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="en_s0", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step (default: the config's BASELINE batch)")
    ap.add_argument("--repeats", type=int, default=5, help="the K timed steps are measured this many times; the median is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from achelous_b200.engine import Engine, compact_spec
    from achelous_b200.nets.Achelous import Achelous
    from achelous_b200.peer_gather import PeerGather
    from achelous_b200.synthetic import make_inputs
    from achelous_b200.weights import fill_state_dict

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the first communicator comes up: fd 1 points at stderr until then, so that
        # stdout carries the JSON line and nothing else
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch or cfg["batch"]
    W = max(args.warmup, 3)
    K = args.steps
    R = max(args.repeats, 1)

    model = Achelous(**cfg["kw"]).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=0), strict=True)
    model = model.to(dev)
    x, xr, pc = make_inputs(B, seed=1234 + rank)
    xd, xrd, pcd = x.to(dev), xr.to(dev), pc.to(dev)
    # with purely random head weights ~1200 of the 2100 anchors per image pass conf >= 0.35 - unrealistically many
    # for the NMS inside the compact plan; shift the objectness bias so that ~120 candidates per image pass (as the golden
    # fixtures do, tests/golden/make_golden.py:calibrate_obj_bias).  Kernel times of the raw forward do not depend on values.
    # Calibrated on the SAME frames on every rank (rank 0's): all ranks must run identical weights.
    cal = [t.to(dev) for t in make_inputs(8, seed=1234)]
    obj_bias, n_cand = calibrate_obj_bias(model, cal[0], cal[1], cal[2], torch)
    comm = torch.cuda.Stream(dev) if world > 1 else None
    # output collection: copy-engine pushes over NVLink peer memory (achelous_b200/peer_gather.py); ACH_BENCH_GATHER=nccl times the
    # NCCL all-gather it replaced (A/B line in profiles/)
    gather_mode = [os.environ.get("ACH_BENCH_GATHER", "peer"), None]

    def barrier():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm)
            dist.barrier()
        torch.cuda.synchronize()

    class Loop:
        """Device-resident step: forward on this rank's B frames [+ ONE all-gather of the packed outputs, in place: the plan writes
        straight into this rank's rows of the gather buffer; two plans / buffers alternate so the gather of step i (side stream)
        overlaps the kernels of step i+1]."""

        def __init__(self, compact):
            ckey = compact_spec() if compact else None
            self.pg = None
            if world == 1:
                self.engines = [Engine(model, B, dev, compact=ckey)]
                self.gathered = None
            else:
                probe = Engine(model, 1, "cpu", dry_run=True, compact=ckey)      # row width / dtype of the packed output
                width, dtype = probe.packed_out.shape[1], probe.packed_out.dtype
                if gather_mode[0] == "peer":
                    try:
                        self.pg = PeerGather(B, width, dtype, dev, copy_streams=int(os.environ.get("ACH_PEER_STREAMS", "2")))
                    except Exception as e:                                        # no IPC / no peer access on this box: say so, use NCCL
                        gather_mode[:] = ["nccl", f"{type(e).__name__}: {e}"[:200]]
                    flag = torch.tensor([1 if self.pg is not None else 0], device=dev)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)                   # all ranks take the same path
                    if flag.item() == 0 and self.pg is not None:
                        self.pg.close()
                        self.pg = None
                        gather_mode[:] = ["nccl", "peer mapping failed on another rank"]
                if self.pg is not None:
                    self.gathered = [self.pg.slot(s_) for s_ in range(2)]
                else:
                    self.gathered = [torch.zeros(world * B, width, device=dev, dtype=dtype) for _ in range(2)]
                self.engines = [Engine(model, B, dev, compact=ckey, out=g[rank * B:(rank + 1) * B]) for g in self.gathered]
                self.done = [None, None]
            for e in self.engines:
                for dst, src in zip(e.input_tensors(), (xd, xrd, pcd)):
                    dst.copy_(src)
                e.forward_static()          # packs weights, captures the CUDA graph
            torch.cuda.synchronize()
            self.i = 0

        def step(self):
            s = self.i % len(self.engines)
            eng = self.engines[s]
            if world > 1:
                main = torch.cuda.current_stream(dev)
                if self.done[s] is not None:
                    main.wait_event(self.done[s])              # the gather that read this buffer two steps ago has finished
                eng.forward_static()
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    if self.pg is not None:
                        seq = self.i + 1
                        self.pg.push(s, seq, comm)             # copy-engine pushes of this rank's rows into every peer's slot
                        self.pg.wait(s, seq, comm)             # ... the step's gather is complete when all ranks' rows are here
                        self.pg.release(s, seq, comm)          # (the consumer of the gathered rows would run before this)
                    else:
                        dist.all_gather_into_tensor(self.gathered[s], eng.packed_out)    # in place: input IS rows [rank*B, (rank+1)*B)
                    self.done[s] = comm.record_event()
            else:
                eng.forward_static()
            self.i += 1

        def close(self):
            self.engines = None
            self.gathered = None
            if self.pg is not None:
                self.pg.close()
                self.pg = None

        def timed(self, sampler=None):
            for _ in range(W):
                self.step()
            barrier()
            if sampler is not None:
                sampler.start()
            reps = []
            ncu_range = os.environ.get("ACH_NCU_RANGE") == "1"   # `ncu --profile-from-start off`: profile exactly the timed steps
            for _ in range(R):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                if ncu_range:
                    torch.cuda.cudart().cudaProfilerStart()
                e0.record()
                for _ in range(K):
                    self.step()
                if world > 1:
                    torch.cuda.current_stream(dev).wait_stream(comm)   # the step's all-gather is part of the step
                e1.record()
                barrier()
                if ncu_range:
                    torch.cuda.cudart().cudaProfilerStop()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                reps.append(t.item())
            clocks = sampler.stop() if sampler is not None else None
            return statistics.median(reps), reps, clocks

    # ---------------- device-resident throughput: raw outputs (headline), then the compact record
    raw = Loop(compact=False)
    eng = raw.engines[0]
    ms_total, reps_raw, clocks = raw.timed(ClockSampler(local_rank) if rank == 0 else None)
    value = world * B * K / (ms_total * 1e-3)

    # on-hardware check of SURVEY.md §8e: gathered rows of every rank == a single-GPU forward of the same frames, bitwise
    gather_check = None
    if world > 1:
        last = (raw.i - 1) % 2
        barrier()
        if rank == 0:
            ok = True
            for r in range(world):
                xs, xrs, pcs = make_inputs(B, seed=1234 + r)
                d, s_, l_, p_ = model(xs[:2].to(dev), xrs[:2].to(dev), pcs[:2].to(dev))
                gd, gs, gl, gp = eng.unpack(raw.gathered[last][r * B:r * B + 2])
                ok &= all(torch.equal(a, b) for a, b in zip(list(d) + [s_, l_, p_], list(gd) + [gs, gl, gp]))
            gather_check = "bitwise-ok" if ok else "MISMATCH"
        barrier()
    # (`raw` stays alive: `eng` below keeps writing into its rows of the peer-mapped gather buffer)
    if os.environ.get("ACH_BENCH_LEGS") == "raw":      # A/B runs (e.g. ACH_BENCH_GATHER=nccl): the device-resident loop only
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms_total / K, "legs": "raw", "gather": gather_mode, "gather_check": gather_check,
                              "repeats": [round(r_ / K, 4) for r_ in reps_raw], "clocks": clocks}), flush=True)
        raw.close()
        if world > 1:
            dist.destroy_process_group()
        return

    cmp_loop = Loop(compact=True)
    ms_c, reps_c, _ = cmp_loop.timed()
    kept = cmp_loop.engines[0].output_views().det_count.float()
    compact = {"value": world * B * K / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / K,
               "nms": {"obj_bias": obj_bias, "candidates_per_frame_mean": round(n_cand, 1), "kept_per_frame_mean": round(kept.mean().item(), 1),
                       "kept_per_frame_max": int(kept.max().item())},
               "bytes_per_frame": cmp_loop.engines[0].packed_out.shape[1], "launches_per_step": len(cmp_loop.engines[0].ops),
               "what": "forward(outputs='compact'): NMS rows (conf 0.35, IoU 0.35, <= 256 per frame), uint8 argmax class maps (320x320) and point "
                       "classes produced inside the launch plan" + (" + in-place all-gather of the compact records" if world > 1 else "")}
    cmp_loop.close()
    del cmp_loop
    torch.cuda.empty_cache()

    # ---------------- end to end through the nn.Module surface with pinned host buffers
    def host_batches(n):
        for _ in range(n):
            yield (xh, xrh, pch)

    def e2e(compact_mode):
        touch = (lambda o: float(o.det_count[0]) + float(o.se_mask[0, 0, 0])) if compact_mode else (lambda o: float(o[3][0, 0, 0]))
        chk = 0.0
        for out in model.stream_forward(host_batches(3), compact=compact_mode):
            chk += touch(out)                        # the caller reads the host result
        reps = []
        for _ in range(R):
            barrier()
            t0 = time.perf_counter()
            n_out = 0
            for out in model.stream_forward(host_batches(K), compact=compact_mode):
                chk += touch(out)
                n_out += 1
            barrier()
            assert n_out == K
            t = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            reps.append(t.item())
        return world * B * K / statistics.median(reps)

    # pinned staging buffers and the copy-driving thread on the GPU's own NUMA node (achelous_b200/hostmem.py); restored afterwards
    # so that the CPU baseline below still sees every host core
    from achelous_b200.hostmem import near_gpu
    numa = {}
    import contextlib
    with (near_gpu(local_rank, numa) if os.environ.get("ACH_BENCH_NUMA", "1") != "0" else contextlib.nullcontext()):
        xh, xrh, pch = x.pin_memory(), xr.pin_memory(), pc.pin_memory()
        h2d = sum(t_.numel() * 4 for t_ in (xh, xrh, pch))
        e2e_compact = e2e(True)
        d2h_compact = B * next(e for k, e in model._engines.items() if k[4] is not None).packed_out.shape[1]
        e2e_raw = e2e(False)
        d2h_raw = B * eng.frame_elems * 4
    model._host_bufs.clear()

    # ---------------- input pre-processing on device: raw camera frames / radar maps / point tables
    pre = None
    if rank == 0:
        from achelous_b200.utils.preprocess import preprocess_image, preprocess_points, preprocess_radar
        g = torch.Generator().manual_seed(7)
        frames = torch.randint(0, 256, (B, 360, 640, 3), dtype=torch.uint8, generator=g).to(dev)
        radar_raw = torch.rand(B, 3, 320, 320, generator=g).to(dev)
        table = torch.randn(300, 5, dtype=torch.float64, generator=g).to(dev)
        idx = torch.randint(0, 300, (B, 512), generator=g).to(dev)

        def pre_step():
            return preprocess_image(frames, (320, 320)), preprocess_radar(radar_raw), preprocess_points(table, idx)
        for _ in range(3):
            pre_step()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(K):
            pre_step()
        p1.record()
        torch.cuda.synchronize()
        pre = {"ms_per_step": p0.elapsed_time(p1) / K, "frames_per_step": B, "launches_per_step": 4,
               "raw_bytes_per_step": frames.numel() + radar_raw.numel() * 4 + idx.numel() * 4,
               "what": "360x640 uint8 RGB frames -> Pillow-exact bicubic letterbox + normalise, radar min-max, point gather + L2 columns"}

    # ---------------- roofline of the dominant kernel, timed live with CUDA events (eager launches)
    roof = measure_dominant(eng, K, torch) if rank == 0 else None

    if rank == 0:
        peaks, which = load_peaks()
        if roof is not None and roof.get("achieved") is not None:
            roof["peak"] = peaks["hbm_gbs"]
            roof["frac"] = roof["achieved"] / peaks["hbm_gbs"]
            roof["peak_source"] = which
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fwd = CpuForward(args.config)
            best, tried = cpu_sweep(fwd, cores, 1, 3, budget_s=25.0)
            cpu = {"value": best[0], "unit": UNIT, "cores": best[2], "kind": fwd.kind,
                   "sample": f"median of 3 forwards per point, best point B={best[1]} x {best[2]} threads of {tried} on {cores} host cores; {fwd.how}"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": cfg["workload"], "config": args.config, "batch_per_gpu": B, "global_batch": B * world,
                           "parallelism": (f"dp{world}: frames sharded, one in-place all-gather of the packed outputs per step ("
                                            + ("copy-engine pushes over NVLink peer memory, achelous_b200/peer_gather.py" if gather_mode[0] == "peer"
                                               else "NCCL" + (f"; peer path unavailable: {gather_mode[1]}" if gather_mode[1] else "")) + ")" if world > 1
                                           else "single GPU"),
                           "outputs": "raw fp32 logits (the reference's return value)",
                           "cuda_graph": True, "l2": "per-step activations (~4 GB) and inputs (158 MB) exceed the 126 MB L2",
                           "weights": "random init (seeded, de-vacuated)"},
                "repeats": {"n": R, "ms_per_step_each": [round(r_ / K, 4) for r_ in reps_raw], "reported": "median"},
                "clocks": clocks,
                "compact": compact,
                "e2e": {"value": e2e_compact, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_compact,
                        "how": "Achelous.stream_forward(compact=True) (public pipelined API): every batch is copied host->device from pinned "
                               "memory, run, and its compact result record (NMS rows + uint8 class maps + point classes: what achelous.py:259-297 "
                               "keeps of the logits) copied device->host and read; copies overlap the neighbouring batches' kernels; wall clock, "
                               "median of the repeats",
                        "raw_logits_value": e2e_raw, "raw_logits_d2h_bytes_per_step": d2h_raw,
                        "raw_logits_how": "stream_forward() with the raw fp32 outputs (4.6 MB per frame) copied out instead",
                        "host_placement": numa},
                "gpu_launches": K * len(eng.ops),
                "launches_per_step": len(eng.ops),
                "roofline": roof, "cpu_baseline": cpu, "preprocess": pre}
        if gather_check is not None:
            line["gather_check"] = gather_check
        print(json.dumps(line), flush=True)
    raw.close()
    if world > 1:
        dist.destroy_process_group()


def calibrate_obj_bias(model, x, xr, pc, torch, conf=0.35, target=120):
    det = model(x, xr, pc)[0]
    flat = torch.cat([d.flatten(2) for d in det], 2)  # (B, 5+K, A)
    cls = torch.sigmoid(flat[:, 5:]).max(1)[0]
    lo, hi = -20.0, 20.0
    for _ in range(30):
        mid = (lo + hi) / 2
        n = ((torch.sigmoid(flat[:, 4] + mid) * cls) >= conf).float().sum(1).mean().item()
        lo, hi = (mid, hi) if n < target else (lo, mid)
    bias = round((lo + hi) / 2, 3)
    with torch.no_grad():
        for k in range(3):
            model.det_head.obj_preds[k].bias += bias      # in-place: the engines repack on their next forward
    n = ((torch.sigmoid(flat[:, 4] + bias) * cls) >= conf).float().sum(1).mean().item()
    return bias, n


def measure_dominant(eng, K, torch):
    """Times every launch once (eager, events), picks the kernel with the largest share and times it again
    K times bracketed by events; algorithmic bytes = unique input + output + weight bytes of that launch."""
    from achelous_b200 import _lib
    stream = torch.cuda.current_stream().cuda_stream
    n = len(eng.ops)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i, (fn, a) in enumerate(eng.ops):
        _lib.check(fn(*a, stream), eng.op_names[i])
        evs[i + 1].record()
    torch.cuda.synchronize()
    t = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    total = sum(t)
    modelled = [i for i in range(n) if eng.algorithmic_bytes(i) is not None]
    top = max(modelled, key=lambda i: t[i])
    fn, a = eng.ops[top]
    nbytes = eng.algorithmic_bytes(top)
    # the launch before it in the plan rewrites its input, so caches are in plan-order state; the tensors are > L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn(*a, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    traffic = None
    # dram__bytes_read+write of this launch from the newest committed `ncu --set full` capture that holds it
    for name in ("r2_final_traffic.json", "r2_traffic.json", "r1s2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f).get(eng.op_names[top], {}).get("dram_bytes")
        except Exception:
            traffic = None
        if traffic is not None:
            break
    # time-weighted roofline fraction of the whole plan: sum(algorithmic bytes) / sum(time) over the modelled launches
    plan_bytes = sum(eng.algorithmic_bytes(i) for i in modelled)
    plan_ms = sum(t[i] for i in modelled)
    return {"bound": "hbm", "kernel": f"{eng.op_names[top]} ({fn.__name__})", "share_of_step": t[top] / total,
            "achieved": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes": nbytes, "launch_ms": ms, "traffic": traffic,
            "plan": {"modelled_launches": len(modelled), "of": n, "algorithmic_bytes": plan_bytes, "eager_ms": plan_ms,
                     "achieved": plan_bytes / (plan_ms * 1e-3) / 1e9},
            "top5": [{"kernel": eng.op_names[i], "ms": round(t[i], 4),
                      "GB/s": round(eng.algorithmic_bytes(i) / (t[i] * 1e-3) / 1e9, 1) if eng.algorithmic_bytes(i) else None}
                     for i in sorted(range(n), key=lambda i: -t[i])[:5]],
            "note": "the top launch by time; every kernel on this path is HBM-bound by arithmetic intensity"}


if __name__ == "__main__":
    main()
