#!/bin/bash
# Final single-GPU visit of the round: parity tests, the driver's bench invocation, launch list, M=64 experiment.
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_r2.sh $TAG
for m64 in 2; do
  PP_M64=$m64 timeout 300 python bench.py --legs value,roofline --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PP_M64=$m64 value', round(d['value']), round(d['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), d['clocks'])"
done
