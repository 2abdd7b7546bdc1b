#!/bin/bash
# tests + default bench + short lines of the other configs.  usage: tools/gpu_quick2.sh <tag>
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | head -20
timeout 600 python bench.py --legs value,e2e,roofline,configs --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; tail -c 300 $OUT/${TAG}_bench_n1.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_n1.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
for k in ("roofline", "roofline_passport_fused", "roofline_wgrad", "roofline_hbm"):
    r = d.get(k)
    print(k, None if not r else (round(r["frac"], 3), r.get("avg_launch_us"), round(r.get("share_of_step", 0), 3)))
for k, v in (d.get("configs") or {}).items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("conv_roofline_frac_whole_step"), v.get("error"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --legs value --no-cpu-baseline --no-graph \
   > $OUT/${TAG}_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches_bench.csv 3 | head -24
