#!/bin/bash
# round-2 final: N = 4 weak-scaling point on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2N4_bench_n4.json 2> gpurun_out/r2N4_bench_n4.err
echo "bench n4 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2N4_bench_n4.json"))
print("n4", d["value"], d["ms_per_step"], d.get("dp_check"), d["e2e"]["value"], d["clocks"])
PY
