#!/bin/bash
# N-GPU visit (gpurun --gpus N): gradient-exchange check, the default workload (BASELINE config 4: V3 + trigger set under
# DDP) with two bucket sizes, and the ImageNet-shaped config (BASELINE config 5).   usage: tools/gpu_n8.sh <tag> <N>
TAG=${1:-r2n8}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/ddp_check.py > $OUT/${TAG}_ddp_check.log 2>&1
echo "ddp_check exit $?"; tail -3 $OUT/${TAG}_ddp_check.log
timeout 400 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline \
    > $OUT/${TAG}_bench_v3_bucket25.json 2> $OUT/${TAG}_bench_v3_bucket25.err
echo "bench v3 (25 MB buckets) exit $?"; head -c 600 $OUT/${TAG}_bench_v3_bucket25.json; echo
PP_BUCKET_MB=8 timeout 400 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline \
    > $OUT/${TAG}_bench_v3_bucket8.json 2> $OUT/${TAG}_bench_v3_bucket8.err
echo "bench v3 (8 MB buckets) exit $?"; head -c 600 $OUT/${TAG}_bench_v3_bucket8.json; echo
PP_BUCKET_MB=8 timeout 400 $RUN --master-port 29514 bench.py --gpus $N --config v1_imagenet --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline \
    > $OUT/${TAG}_bench_imagenet.json 2> $OUT/${TAG}_bench_imagenet.err
echo "bench imagenet exit $?"; head -c 600 $OUT/${TAG}_bench_imagenet.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline > $OUT/${TAG}_bench_v3_n1.json 2> $OUT/${TAG}_bench_v3_n1.err
echo "bench v3 N=1 exit $?"; head -c 400 $OUT/${TAG}_bench_v3_n1.json; echo
timeout 300 python bench.py --config v1_imagenet --steps 20 --warmup 5 --legs value,e2e --no-cpu-baseline > $OUT/${TAG}_bench_imagenet_n1.json 2> $OUT/${TAG}_bench_imagenet_n1.err
echo "bench imagenet N=1 exit $?"; head -c 400 $OUT/${TAG}_bench_imagenet_n1.json; echo
