#!/bin/bash
# A/B of the side-stream weight gradients: tests, then value / e2e with and without the overlap (graph replay and eager).
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -2
grep -E "^(FAILED|ERROR)|^E  " $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | head -10
for ov in 0 1; do for gr in "" "--no-graph"; do
  PP_NO_WGRAD_OVERLAP=$ov timeout 300 python bench.py --legs value,e2e --no-cpu-baseline $gr 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_overlap=$ov', '$gr' or 'graph', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done; done
