#!/bin/bash
# round-2 final call C (8 GPUs): weak-scaling bench + --check-dp on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 10 240 $TR --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2C_bench_n8.json 2> gpurun_out/r2C_bench_n8.err
echo "bench n8 rc=$?"
timeout -k 10 240 $TR --master-port 29542 bench.py --gpus 8 --check-dp > gpurun_out/r2C_check_dp_n8.json 2> gpurun_out/r2C_check_dp_n8.err
echo "check-dp n8 rc=$?"
cat gpurun_out/r2C_check_dp_n8.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2C_bench_n8.json"))
print("n8", d["value"], d["ms_per_step"], d.get("dp_check"), d["e2e"]["value"], d["clocks"])
PY
tail -n 3 gpurun_out/r2C_bench_n8.err
