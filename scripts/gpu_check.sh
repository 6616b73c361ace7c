#!/bin/bash
# One GPU-box visit: parity tests (each family in its own process so a faulting kernel cannot poison the rest),
# a short bench, and the ncu captures that back the roofline numbers.  Usage: gpurun -- bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_streaming.py tests/test_gpu_widened.py -m gpu -q --timeout 300 > $OUT/pytest_streaming_$TAG.log 2>&1
echo "streaming exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 300 python -m pytest tests/test_gpu_contrastive.py -m gpu -q --timeout 120 -k "similarity" > $OUT/pytest_similarity_$TAG.log 2>&1
echo "similarity exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 300 python -m pytest tests/test_gpu_contrastive.py -m gpu -q --timeout 120 -k "pair_kernel_sees" > $OUT/pytest_pair_$TAG.log 2>&1
echo "pair exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_contrastive.py -m gpu -q --timeout 300 -k "not similarity and not pair_kernel_sees" > $OUT/pytest_contrastive_$TAG.log 2>&1
echo "contrastive exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/summary_$TAG.txt
tail -c 3000 $OUT/bench_$TAG.json
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
  timeout 600 $CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches exit $?" | tee -a $OUT/summary_$TAG.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower_stream|mse_stream|attn_kl" -s 6 -c 9 -o $OUT/prof_mse_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?" | tee -a $OUT/summary_$TAG.txt
fi
for f in $OUT/pytest_*_$TAG.log; do echo "== $f"; tail -n 25 $f; done
cat $OUT/smoke_$TAG.log | tail -5
