#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LAV_BENCH_GEMM_OUT=r2x_gemm_head.json timeout 300 python tools/bench_gemm.py --head 2>&1 | grep tag
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_rowops_gpu.py tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_multitask_gpu.py tests/test_dropin.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1
tail -n 6 gpurun_out/r2x_tests.log
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'], {k:v['ms'] for k,v in d['kernels'].items()})"
LAV_MERGE_HEADS=0 timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2x_bench_nomerge.json 2> gpurun_out/r2x_bench_nomerge.err
python -c "
import json
d=json.load(open('gpurun_out/r2x_bench_nomerge.json'))
print('no merge:', d['value'], d['ms_per_step'], d['loss'])"
