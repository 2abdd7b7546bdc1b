"""The benchmark and the driver entry points must keep working: run bench.py end to end at a tiny batch (all
legs) and __graft_entry__.smoke() on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_bench_contract_small_batch():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", "32", "--steps", "3", "--warmup", "3",
                        "--cpu-steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-1000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 32 * 3 * 32 * 32 * 4 + 32 * 8
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1.5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"].startswith("ResNet18 V2")


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port"


def test_graft_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.smoke()
