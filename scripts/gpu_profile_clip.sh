#!/bin/bash
# ncu evidence for the fused contrastive kernels.  Usage: gpurun -- bash scripts/gpu_profile_clip.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload lclip --steps 3 --warmup 3"
timeout 300 $CMD > $OUT/clip_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/clip_launches_$TAG.csv $CMD > $OUT/clip_ncu_launches_$TAG.log 2>&1
echo "launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"clip_fwd_kernel|clip_bwd_pair|clip_gt_gemm" -s 3 -c 3 -o $OUT/prof_clip_$TAG -f $CMD > $OUT/clip_ncu_full_$TAG.log 2>&1
echo "full exit $?"
CMD2="python bench.py --workload sweep --steps 3 --warmup 3"
timeout 300 $CMD2 > $OUT/sweep_plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"clip_" -c 60 --csv --log-file $OUT/sweep_launches_$TAG.csv $CMD2 > $OUT/sweep_ncu_launches_$TAG.log 2>&1
echo "sweep launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"clip_fwd_kernel|clip_bwd_pair|clip_gt_gemm" -s 3 -c 3 -o $OUT/prof_sweep_$TAG -f $CMD2 > $OUT/sweep_ncu_full_$TAG.log 2>&1
echo "sweep full exit $?"
tail -2 $OUT/clip_plain_$TAG.log | cut -c1-600
