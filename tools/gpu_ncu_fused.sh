#!/bin/bash
# ncu --set full of the single-kernel passport block (layer4 geometry, per-GPU batch 1024 + 2) and, for comparison,
# of the kernel sequence it replaces; CUDA-event A/B timing of both.  usage: tools/gpu_ncu_fused.sh <tag>
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
for L in layer4 layer4s2 layer4sc; do
  python tools/profile_layer.py --layer $L --batch 1026 --iters 30 --warmup 5 >> $OUT/${TAG}_ab_fused.txt 2>&1
  python tools/profile_layer.py --layer $L --batch 1026 --iters 30 --warmup 5 --no-fused >> $OUT/${TAG}_ab_sequence.txt 2>&1
done
cat $OUT/${TAG}_ab_fused.txt $OUT/${TAG}_ab_sequence.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'passport_fused_kernel' -c 2 -f \
   -o $OUT/${TAG}_ncu_fused python tools/profile_layer.py --layer layer4 --batch 1026 --iters 1 --warmup 1 \
   > $OUT/${TAG}_ncu_fused.log 2>&1
echo "ncu fused exit $?"
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'tapgemm_kernel|bn_finalize_kernel|affine_apply_kernel|passport_gemv_kernel|sign_loss_kernel' -c 8 -f \
   -o $OUT/${TAG}_ncu_sequence python tools/profile_layer.py --layer layer4 --batch 1026 --iters 1 --warmup 0 --no-fused \
   > $OUT/${TAG}_ncu_sequence.log 2>&1
echo "ncu sequence exit $?"
ls -la $OUT/${TAG}_*
