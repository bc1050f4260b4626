#!/bin/bash
# round-2 call H2 (2 GPUs): data-parallel correctness (--check-dp), bit-identical replicas, overlapped all-reduce
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29511 bench.py --gpus 2 --check-dp > gpurun_out/r2h_check_dp.json 2> gpurun_out/r2h_check_dp.err
echo "check-dp rc=$?"
timeout -k 10 300 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "bench n2 rc=$?"
cat gpurun_out/r2h_check_dp.json; tail -n 5 gpurun_out/r2h_check_dp.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench_n2.json"))
print("n2", d["value"], d["ms_per_step"], d.get("dp_check"), d["e2e"]["value"])
PY
