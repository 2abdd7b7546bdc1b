#!/bin/bash
# The driver's own invocation at N GPUs (default flags: graph replay, all legs), plus --no-graph for comparison.
TAG=${1:-r2n8g}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err
echo "N=$N default exit $?"; tail -2 $OUT/${TAG}_bench_graph.err | cut -c1-200
timeout 400 $RUN --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --legs value,e2e --no-graph > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err
echo "N=$N eager exit $?"
timeout 300 $RUN --master-port 29523 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "N=$N reference arm exit $?"; head -c 400 $OUT/${TAG}_bench_reference.json; echo
python - <<PY
import json
for mode in ("graph", "eager"):
    try:
        d = json.loads([l for l in open("$OUT/${TAG}_bench_%s.json" % mode) if l.startswith("{")][-1])
        print($N, mode, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
              "launches", d["gpu_launches"], "graph", d["cuda_graph"], "clocks", d["clocks"])
    except Exception as e:
        print($N, mode, "failed", e)
PY
