#!/bin/bash
# round-2 final call B (2 GPUs): weak-scaling bench + --check-dp on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2B_bench_n2.json 2> gpurun_out/r2B_bench_n2.err
echo "bench n2 rc=$?"
timeout -k 10 300 $TR --master-port 29532 bench.py --gpus 2 --check-dp > gpurun_out/r2B_check_dp_n2.json 2> gpurun_out/r2B_check_dp_n2.err
echo "check-dp rc=$?"
timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2B_bench_n1.json 2> gpurun_out/r2B_bench_n1.err
cat gpurun_out/r2B_check_dp_n2.json
for f in n2 n1; do python - <<PY
import json
d=json.load(open("gpurun_out/r2B_bench_$f.json"))
print("$f", d["value"], d["ms_per_step"], d.get("dp_check"), d["e2e"]["value"])
PY
done
