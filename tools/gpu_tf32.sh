#!/bin/bash
# GPU-box visit for the TF32 configuration (BASELINE config 2): parity tests, bench line, launch list.
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tf32.py -m gpu -q > $OUT/${TAG}_pytest_tf32.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest_tf32.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " $OUT/${TAG}_pytest_tf32.log | head -30
timeout 600 python bench.py --config v1_alexnet --legs value,e2e,roofline --no-cpu-baseline > $OUT/${TAG}_bench_alexnet_tf32.json 2> $OUT/${TAG}_bench_alexnet_tf32.err
echo "bench tf32 exit $?"; tail -c 400 $OUT/${TAG}_bench_alexnet_tf32.err; head -c 1500 $OUT/${TAG}_bench_alexnet_tf32.json; echo
timeout 600 python bench.py --config v1_alexnet_bf16 --legs value --no-cpu-baseline > $OUT/${TAG}_bench_alexnet_bf16.json 2> $OUT/${TAG}_bench_alexnet_bf16.err
echo "bench bf16 exit $?"; head -c 400 $OUT/${TAG}_bench_alexnet_bf16.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
   --log-file $OUT/${TAG}_launches_alexnet_tf32.csv python bench.py --config v1_alexnet --steps 2 --warmup 1 --legs value --no-cpu-baseline --no-graph \
   > $OUT/${TAG}_launches_alexnet_tf32.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py $OUT/${TAG}_launches_alexnet_tf32.csv 3 | head -30
