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
                        "--cpu-steps", "1", "--legs", "value,e2e,roofline,eager,dropin,shared"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-1000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3
    # V3 default: 32 images + 2 trigger images per step, all copied from pinned host memory in the e2e leg
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 34 * 3 * 32 * 32 * 4 + 34 * 8
    assert d["e2e"]["d2h_bytes_per_step"] == 16 and d["last_step"]["metric_reads"] == 3
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1.5
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"].startswith("ResNet18 V3") and d["config"]["trigger_images_per_step_per_gpu"] == 2
    # the unmodified reference on this GPU, and its own trainer driving the patched blocks
    assert d["torch_eager_gpu"]["fp32_as_shipped"]["value"] > 0
    assert d["reference_trainer_on_patched_blocks"]["value"] > 0
    assert d["reference_trainer_on_patched_blocks"]["library_launches"] > 100


def test_bench_other_configs_and_graph_mode():
    """BASELINE configs 2, 3 and 5 through --config, and the CUDA-graph step through --graph, at tiny batches."""
    for extra in (["--config", "v1_alexnet"], ["--config", "v2_cifar100"], ["--config", "v1_imagenet", "--batch", "4"],
                  ["--no-graph"]):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", "16", "--steps", "2", "--warmup",
                            "3", "--no-cpu-baseline", "--legs", "value,e2e", *extra], capture_output=True, text=True,
                           timeout=600, cwd=ROOT)
        assert p.returncode == 0, (extra, p.stderr[-2000:])
        d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
        assert d["value"] > 0 and d["e2e"]["value"] > 0, extra
        assert d["cuda_graph"] == ("--no-graph" not in extra)


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("ResNet18 V3")


def test_graft_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.smoke()
