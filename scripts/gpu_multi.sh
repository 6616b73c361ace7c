#!/bin/bash
# N-GPU visit: NCCL parity tests + the sharded bench.  Usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N [tag] [skip_tests]
N=${1:-2}
TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$3" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 400 -x > $OUT/pytest_multi_$TAG.log 2>&1
echo "multi exit $?" | tee -a $OUT/summary_$TAG.txt
tail -n 30 $OUT/pytest_multi_$TAG.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench exit $?" | tee -a $OUT/summary_$TAG.txt
tail -c 3000 $OUT/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("bench:", d["n_gpus"], d["value"], d["ms_per_step"], "launches", d["gpu_launches"], d["config"]["timing"], "parity", d["parity"]["ok"], json.dumps(d["parity"]["routes"]), d["parity"]["grad_rel_l2_api_bf16"])
    print("roofline:", json.dumps(d["roofline"]["kernels"]), json.dumps(d["roofline"]["step"]))
    print("e2e:", d["e2e"])
    for k, v in d.get("stages", {}).items():
        print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], v["parity"]["ok"])
    if "lclip" in d:
        l = d["lclip"]; print("lclip", l["value"], l["ms_per_step"], l["timing"], l["roofline"]["step"], l["parity"]["ok"], json.dumps(l["parity"]["routes"]), l["gpu_launches"])
except Exception as e:
    print("bench parse failed:", e)
PY
