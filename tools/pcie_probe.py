"""Host<->device copy bandwidth of the box (pinned memory), to interpret the e2e number."""
import time
import torch
for mb in (16, 158, 296):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n).pin_memory()
    d = torch.empty(n, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb} MB pinned: {dt*1e3:.2f} ms  {mb/1024/dt:.1f} GiB/s")
hp = torch.empty(158 * 1024 * 1024 // 4)
t0 = time.perf_counter(); d2 = hp.cuda(); torch.cuda.synchronize(); print("pageable H2D 158MB", (time.perf_counter()-t0)*1e3, "ms")
t0 = time.perf_counter(); hp.pin_memory(); print("pin_memory() of 158MB", (time.perf_counter()-t0)*1e3, "ms")
