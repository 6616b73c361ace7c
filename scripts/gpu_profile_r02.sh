#!/bin/bash
# Round-2 evidence on one GPU: default bench, reference arm, launch list and one ncu --set full capture of the three
# tcgen05 kernels at the sweep size.  Usage: gpurun -- bash scripts/gpu_profile_r02.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref_n1.json 2> $OUT/${TAG}_bench_ref_n1.err
echo "reference exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --no-parity"
timeout 600 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_sweep.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"clip_fwd_kernel|clip_bwd_pair_kernel|clip_g_tiles_kernel|clip_gt_gemm_kernel|clip_post|clip_prep|clip_finish2" -c 8 -o $OUT/${TAG}_prof_sweep -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
# the reports are large (gpurun brings back at most 64 MiB): extract the counters here and keep only the CSV
python scripts/ncu_extract.py $OUT/${TAG}_prof_sweep.ncu-rep $OUT/${TAG}_ncu_full_clip_sweep.csv && rm -f $OUT/${TAG}_prof_sweep.ncu-rep
tail -c 600 $OUT/${TAG}_bench_n1.err
