#!/bin/bash
# Full GPU-box visit: parity tests, bench line (all legs), launch list, A/B + ncu of the single-kernel passport block,
# ncu --set full of the dominant conv kernel (traffic for the roofline record).   usage: tools/gpu_visit.sh <tag>
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_r2.sh $TAG
for Lr in layer4 layer4s2 layer4sc; do
  python tools/profile_layer.py --layer $Lr --batch 1026 --iters 30 --warmup 5 >> $OUT/${TAG}_ab_fused.txt 2>&1
done
cat $OUT/${TAG}_ab_fused.txt | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'passport_fused_kernel' -c 1 -f \
   -o $OUT/${TAG}_ncu_fused python tools/profile_layer.py --layer layer4 --batch 1026 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_fused.log 2>&1
echo "ncu fused exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pxn_kernel|affine_apply_kernel|column_reduce_kernel|bwd_dz_kernel' -c 10 -f \
   -o $OUT/${TAG}_ncu_pxn_layer1 python tools/profile_layer.py --layer layer1 --kind conv --batch 1026 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_pxn_layer1.log 2>&1
echo "ncu pxn exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
   --log-file $OUT/${TAG}_launches_imagenet.csv python bench.py --config v1_imagenet --steps 2 --warmup 1 --legs value --no-cpu-baseline --no-graph \
   > $OUT/${TAG}_launches_imagenet.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches_imagenet.csv 3 | head -24
ls -la $OUT/${TAG}_* | head -30
