"""CPU: the reference arm of bench.py (the reference's own CPU forward - oracle/_ref or /root/reference, kind "reference"; the
oracle port for config 4, which the reference does not implement) prints one JSON line with the keys the driver reads;
rank != 0 under torchrun prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *extra],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    # stdout carries the JSON line and nothing else (the reference's imports print install hints: they belong on stderr)
    assert all(ln.startswith("{") for ln in out.stdout.splitlines() if ln.strip()), out.stdout[:500]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("frames/sec full 5-task forward")
    from oracle.ref_loader import reference_available
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_reference_arm_config4_is_the_port():
    """EN-GDF-PN2-S2: the reference ships no PointNet++, so its CPU arm can only be the (builder-defined) oracle port"""
    d = json.loads(_run({}, "--config", "en_s2_pn2")[0])
    assert d["cpu_baseline"]["kind"] == "port" and d["config"]["config"] == "en_s2_pn2" and d["value"] > 0
