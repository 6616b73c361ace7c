#!/bin/bash
# Quick counters (duration, cycles, tensor-pipe activity, instructions) of the contrastive kernels of one sweep step.
# Usage: gpurun -- bash scripts/ncu_quick.sh [kernel-regex] [tag]
RE=${1:-clip_fwd_kernel|clip_g_tiles_kernel|clip_gt_gemm_kernel}
TAG=${2:-quick}
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"$RE" -c 4 --csv --log-file gpurun_out/ncu_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/ncu_$TAG.csv")) if len(r) > 10]
h = rows[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iid = h.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((r[iid], r[ik][:44]), {})[r[im]] = r[iv]
for k, v in out.items():
    print(k, {a.split(".")[0][-22:]: b for a, b in v.items()})
PY
