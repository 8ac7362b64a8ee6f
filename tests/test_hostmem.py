"""CPU: achelous_b200.hostmem.near_gpu is best effort - without a GPU / NVML it must leave the process affinity alone, say why,
and never raise; with a fake NVML answer it narrows the affinity inside the block and restores it afterwards."""
import os

from achelous_b200 import hostmem


def test_near_gpu_without_nvml_is_a_noop():
    before = os.sched_getaffinity(0)
    info = {}
    with hostmem.near_gpu(0, info):
        assert os.sched_getaffinity(0) == before
    assert os.sched_getaffinity(0) == before
    assert info["cpus_before"] == len(before) and info["cpus_bound"] is None and info["note"]


def test_near_gpu_binds_and_restores(monkeypatch):
    before = os.sched_getaffinity(0)
    if len(before) < 2:
        import pytest
        pytest.skip("needs two CPUs in the cpuset")
    local = set(sorted(before)[: max(1, len(before) // 2)])
    monkeypatch.setattr(hostmem, "gpu_local_cpus", lambda index: local | {10 ** 6})    # CPUs outside the cpuset are ignored
    info = {}
    with hostmem.near_gpu(0, info):
        assert os.sched_getaffinity(0) == local
    assert os.sched_getaffinity(0) == before
    assert info["cpus_bound"] == len(local) and info["note"] is None


def test_near_gpu_restores_after_an_exception(monkeypatch):
    before = os.sched_getaffinity(0)
    if len(before) < 2:
        import pytest
        pytest.skip("needs two CPUs in the cpuset")
    monkeypatch.setattr(hostmem, "gpu_local_cpus", lambda index: {min(before)})
    try:
        with hostmem.near_gpu(0):
            raise RuntimeError("boom")
    except RuntimeError:
        pass
    assert os.sched_getaffinity(0) == before
