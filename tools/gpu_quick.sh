#!/bin/bash
# Lean GPU-box visit: parity tests + bench line + launch list.  usage: tools/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n1.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "hbm", d["roofline_hbm"]["frac"] if d.get("roofline_hbm") else None)
except Exception as e:
    print("bench parse failed", e)
PY
if [ -n "$NCU_STEM" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_|wgrad_finalize' -c 4 -f \
   -o $OUT/${TAG}_ncu_stem python tools/profile_layer.py --layer stem --kind conv --batch 1184 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_stem.log 2>&1
echo "ncu stem exit $?"
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --legs value --no-cpu-baseline \
   > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py $OUT/${TAG}_launches_bench.csv 3 | head -24
