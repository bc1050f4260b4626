#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 1 0 1 0; do
LAV_LN_STAGED=$v timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2D_bench_staged$v.json 2> gpurun_out/r2D_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2D_bench_staged$v.json'))
print('LN_STAGED=$v', d['value'], d['ms_per_step'], d['kernels']['layernorm_bwd'])"
done
