"""2-GPU probe: where does a forward on cuda:1 fail (single thread, then nn.DataParallel)?"""
import ctypes
import os
import sys
import threading
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200.nets.Achelous import Achelous  # noqa: E402
from achelous_b200.synthetic import make_inputs  # noqa: E402
from achelous_b200.weights import fill_state_dict  # noqa: E402

rt = ctypes.CDLL("libcudart.so.12")


def real_device():
    d = ctypes.c_int(-1)
    rt.cudaGetDevice(ctypes.byref(d))
    return d.value


kw = dict(num_det=7, num_seg=9, phi="S0", resolution=320, backbone="en", neck="gdf", pc_seg="pn", pc_channels=5, pc_classes=8,
          nano_head=True, spp=True)
model = Achelous(**kw).eval()
model.load_state_dict(fill_state_dict(model.state_dict(), seed=2))
x, xr, pc = make_inputs(2, seed=1)
for graph in (False, True):
    model.use_cuda_graph = graph
    try:
        m1 = model.to("cuda:1")
        print("single thread, module on cuda:1, torch current", torch.cuda.current_device(), "real", real_device())
        o = m1(x.cuda(1), xr.cuda(1), pc.cuda(1))
        torch.cuda.synchronize(1)
        print("  ok graph=%s" % graph, o[1].device, float(o[1].abs().mean()), "real after", real_device())
    except Exception:
        traceback.print_exc()
model = model.to("cuda:0")
model.use_cuda_graph = False
ref = model(x.cuda(0), xr.cuda(0), pc.cuda(0))
print("cuda:0 ok; now a replica-style call on device 1 from a worker thread")


def worker():
    try:
        with torch.cuda.device(1), torch.cuda.stream(torch.cuda.current_stream(1)):
            print("  worker torch current", torch.cuda.current_device(), "real", real_device())
            rep = model._replicate_for_data_parallel()
            o = rep(x.cuda(1), xr, pc.cuda(1))
            torch.cuda.synchronize(1)
            print("  worker ok", o[1].device, torch.equal(o[1].cpu(), ref[1].cpu()))
    except Exception:
        traceback.print_exc()


t = threading.Thread(target=worker)
t.start()
t.join()
for graph in (False, True):
    model.use_cuda_graph = graph
    model._engines.clear()
    try:
        dp = torch.nn.DataParallel(model)
        x4, xr4, pc4 = make_inputs(4, seed=1)
        o = dp(x4.cuda(0), xr4, pc4.cuda(0))
        print("DataParallel ok graph=%s" % graph, o[1].shape, o[1].device)
    except Exception as e:
        print("DataParallel FAILED graph=%s:" % graph, str(e)[-300:])
