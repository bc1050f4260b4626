#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rowops_gpu.py -m gpu -x -q > gpurun_out/r2s_rowops_tests.log 2>&1
tail -n 15 gpurun_out/r2s_rowops_tests.log
timeout 300 python tools/bench_rowops.py gpurun_out/r2s_rowops.json > gpurun_out/r2s_rowops.log 2>&1
cat gpurun_out/r2s_rowops.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_dropout_gpu.py tests/test_graph_gpu.py -m gpu -x -q > gpurun_out/r2s_model_tests.log 2>&1
tail -n 5 gpurun_out/r2s_model_tests.log
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], {k:v['ms'] for k,v in d['kernels'].items()})"
