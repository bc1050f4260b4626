#!/bin/bash
# round-2 call H (2 GPUs): data-parallel correctness (--check-dp) and the overlapped, graph-captured gradient all-reduce
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29511 bench.py --gpus 2 --check-dp > gpurun_out/r2h_check_dp.json 2> gpurun_out/r2h_check_dp.err
echo "check-dp rc=$?"
timeout -k 10 400 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "bench n2 rc=$?"
LAV_GRAPH_NCCL=0 timeout -k 10 400 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2_nooverlap.json 2> gpurun_out/r2h_bench_n2_nooverlap.err
echo "bench n2 (LAV_GRAPH_NCCL=0) rc=$?"
timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
cat gpurun_out/r2h_check_dp.json; tail -n 5 gpurun_out/r2h_check_dp.err
for f in n2 n2_nooverlap n1; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h_bench_$f.json"))
    print("$f", d["value"], d["ms_per_step"], d.get("dp_check"), d["e2e"]["value"])
except Exception as e:
    print("$f failed", e)
PY
done
tail -n 5 gpurun_out/r2h_bench_n2.err
