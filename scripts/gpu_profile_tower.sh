#!/bin/bash
# ncu evidence for the streaming (tower) path.  Usage: gpurun -- bash scripts/gpu_profile_tower.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu"
timeout 300 $CMD > $OUT/tower_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/tower_launches_$TAG.csv $CMD > $OUT/tower_ncu_launches_$TAG.log 2>&1
echo "launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower_stream" -s 2 -c 3 -o $OUT/prof_tower_$TAG -f $CMD > $OUT/tower_ncu_full_$TAG.log 2>&1
echo "full exit $?"
tail -1 $OUT/tower_plain_$TAG.log | cut -c1-400
