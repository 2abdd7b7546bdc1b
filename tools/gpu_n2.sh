#!/bin/bash
# 2-GPU visit (gpurun --gpus 2): parity tests on one GPU, gradient-exchange check and a short bench on two.
TAG=${1:-n2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/ddp_check.py > $OUT/${TAG}_ddp_check.log 2>&1
echo "ddp_check exit $?"; tail -3 $OUT/${TAG}_ddp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --legs value,e2e > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
echo "bench n2 exit $?"; tail -c 1500 $OUT/${TAG}_bench_n2.json; tail -3 $OUT/${TAG}_bench_n2.err
