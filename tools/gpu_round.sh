#!/bin/bash
# One GPU-box visit: parity tests, bench line, launch list, ncu --set full captures.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh [tag]      (run through gpurun from the repo root)
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== bench"
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; tail -c 3000 $OUT/${TAG}_bench_n1.json
echo "== ncu full: layer4 passport block at batch 1184"
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'tapgemm_kernel|wgrad_kernel|affine_apply_kernel|column_reduce_kernel|bwd_dz_kernel' -c 7 -f \
   -o $OUT/${TAG}_ncu_layer4 python tools/profile_layer.py --layer layer4 --batch 1184 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_layer4.log 2>&1
echo "ncu layer4 exit $?"
echo "== ncu full: layer1 conv block pointwise passes at batch 1184 (HBM-bound kernels)"
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'affine_apply_kernel|column_reduce_kernel|bwd_dz_kernel' -c 3 -f \
   -o $OUT/${TAG}_ncu_layer1_pointwise python tools/profile_layer.py --layer layer1 --kind conv --batch 1184 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_layer1_pointwise.log 2>&1
echo "ncu layer1 exit $?"
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --legs value --no-cpu-baseline \
   > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
ls -la $OUT | head -40
