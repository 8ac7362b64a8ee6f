"""Host-side placement for the pinned staging buffers of `Achelous.stream_forward` on multi-socket boxes.

Pinned pages are allocated on the NUMA node of the thread that asks for them.  With one process per GPU and no placement, the
staging buffers of the GPUs behind the other socket sit on the wrong node and every host<->device copy crosses the socket
interconnect (8-GPU box: 65 k frames/s end to end against 116 k device-resident, profiles/r2_bench_8gpu.json).  `near_gpu`
restricts the calling process to the CPUs NVML reports as local to the GPU for the duration of a `with` block: allocate the
pinned buffers and drive the copies inside it."""
import contextlib
import os


def gpu_local_cpus(index):
    """CPUs local to CUDA device `index` (NVML's ideal affinity, matched by PCI bus id so that CUDA_VISIBLE_DEVICES does not matter)"""
    import pynvml
    import torch
    pynvml.nvmlInit()
    pr = torch.cuda.get_device_properties(index)
    bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
    words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
    return {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}


@contextlib.contextmanager
def near_gpu(index, info=None):
    """Runs the block with the process affinity narrowed to the GPU's local CPUs (best effort: a box without NVML / NUMA
    information, or a cpuset that excludes those CPUs, leaves the affinity as it is).  `info`, a dict, receives what was done."""
    old = os.sched_getaffinity(0)
    new, why = None, None
    try:
        cpus = gpu_local_cpus(index) & old
        if cpus and cpus != old:
            new = cpus
        else:
            why = "GPU-local CPUs == current affinity" if cpus else "GPU-local CPUs outside this process's cpuset"
    except Exception as e:   # noqa: BLE001 - placement is an optimisation, never a failure
        why = f"{type(e).__name__}: {e}"[:120]
    if info is not None:
        info.update({"cpus_before": len(old), "cpus_bound": len(new) if new else None, "note": why})
    if new:
        os.sched_setaffinity(0, new)
    try:
        yield
    finally:
        if new:
            os.sched_setaffinity(0, old)
