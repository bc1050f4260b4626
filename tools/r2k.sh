#!/bin/bash
# round-2 call K: the full training step with the CTA-pair GEMM path (6-stage ring, elected arrives) vs the 1-CTA path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LAV_GEMM_PAIR=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench_pair.json 2> gpurun_out/r2k_bench_pair.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench_single.json 2> gpurun_out/r2k_bench_single.err
for f in pair single; do python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench_$f.json"))
print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["kernels"]["gemm"])
PY
done
