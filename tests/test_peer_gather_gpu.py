"""GPU (>= 2 devices): the copy-engine all-gather over NVLink peer memory (achelous_b200/peer_gather.py) delivers exactly what
NCCL's all-gather delivers, over several steps of the two-slot flag protocol (tools/peer_probe.py under torchrun)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_gather_equals_nccl_all_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2; log in profiles/r2_peer_gather_2gpu.log)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "peer_probe.py"), "--mb", "32", "--steps", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["check"] == "ok" and line["world"] == 2
