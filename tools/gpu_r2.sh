#!/bin/bash
# Round-2 GPU-box visit: parity tests, bench line (all legs), launch list.  usage: tools/gpu_r2.sh <tag> [pytest-args]
TAG=${1:-r2a}
shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q "$@" > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -5
grep -E "^(FAILED|ERROR)" $OUT/${TAG}_pytest_gpu.log | head -30
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; tail -c 600 $OUT/${TAG}_bench_n1.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n1.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k in ("roofline", "roofline_passport_fused", "roofline_wgrad", "roofline_hbm"):
        r = d.get(k)
        print(k, None if not r else (round(r["frac"], 3), round(r["avg_launch_us"], 1) if "avg_launch_us" in r else None, round(r.get("share_of_step", 0), 3)))
    print("eager", d.get("torch_eager_gpu")); print("dropin", d.get("reference_trainer_on_patched_blocks"))
    print("small", d.get("small_batch")); print("configs", d.get("configs")); print("cpu", d.get("cpu_baseline"))
    print("shared", d.get("value_shared_trunk")); print("clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
except Exception as e:
    print("bench parse failed", e)
PY
fi
if [ -z "$SKIP_LAUNCHES" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --legs value --no-cpu-baseline --no-graph \
   > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py $OUT/${TAG}_launches_bench.csv 3 | head -40
fi
