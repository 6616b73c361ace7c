#!/bin/bash
# pipeline tests + default bench on one GPU.  Usage: gpurun -- bash scripts/gpu_check2.sh [tag]
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -x > $OUT/pytest_pipeline_$TAG.log 2>&1
echo "pipeline exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_contrastive.py tests/test_gpu_fullsize.py tests/test_gpu_widened.py -m gpu -q --timeout 600 > $OUT/pytest_contrastive_$TAG.log 2>&1
echo "contrastive+fullsize exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/summary_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/summary_$TAG.txt
for f in $OUT/pytest_*_$TAG.log; do echo "== $f"; tail -n 30 $f; done
tail -5 $OUT/smoke_$TAG.log
tail -c 2500 $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("bench:", d["value"], d["ms_per_step"], "launches", d["gpu_launches"], "parity", d["parity"]["ok"], json.dumps(d["parity"]["routes"]), d["parity"]["grad_rel_l2_api_bf16"])
    print("roofline:", json.dumps(d["roofline"]["kernels"]), json.dumps(d["roofline"]["step"]))
    print("e2e:", d["e2e"], "cpu:", d.get("cpu_baseline"))
    for k, v in d.get("stages", {}).items():
        print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], {a: b["frac"] for a, b in v["roofline"]["kernels"].items()}, v["parity"])
    if "lclip" in d:
        l = d["lclip"]; print("lclip", l["value"], l["ms_per_step"], l["roofline"]["step"], l["roofline"]["kernels"], l["parity"]["ok"], l["gpu_launches"])
except Exception as e:
    print("bench parse failed:", e)
PY
