#!/bin/bash
# CUDA-graph step under DDP (NCCL all-reduces captured with the step): default (graph) vs --no-graph at N GPUs, and
# the full default bench at N=1.   usage: tools/gpu_graph_ddp.sh <tag> <N>
TAG=${1:-r2g2}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in graph eager; do
  extra=""; [ $mode = eager ] && extra="--no-graph"
  timeout 300 $RUN --master-port 2951$((RANDOM % 9)) bench.py --gpus $N --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline $extra > $OUT/${TAG}_n${N}_$mode.json 2> $OUT/${TAG}_n${N}_$mode.err
  echo "N=$N $mode exit $?"; tail -2 $OUT/${TAG}_n${N}_$mode.err | cut -c1-200
done
python - <<PY
import json
for mode in ("eager", "graph"):
    try:
        d = json.loads([l for l in open("$OUT/${TAG}_n${N}_%s.json" % mode) if l.startswith("{")][-1])
        print($N, mode, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
              "launches", d["gpu_launches"], "loss", d.get("last_step", {}).get("loss"), "graph", d["cuda_graph"])
    except Exception as e:
        print($N, mode, "failed", e)
PY
