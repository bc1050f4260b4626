#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rowops_gpu.py tests/test_input_pipeline.py -m gpu -x -q > gpurun_out/r2u_rowops_tests.log 2>&1
tail -n 8 gpurun_out/r2u_rowops_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_multitask_gpu.py -m gpu -x -q > gpurun_out/r2u_model_tests.log 2>&1
tail -n 5 gpurun_out/r2u_model_tests.log
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2u_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('input_pipeline'), {k:v['ms'] for k,v in d['kernels'].items()})"
LAV_FUSE_CASTS=0 timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2u_bench_nofuse.json 2> gpurun_out/r2u_bench_nofuse.err
python -c "
import json
d=json.load(open('gpurun_out/r2u_bench_nofuse.json'))
print('no fuse:', d['value'], d['ms_per_step'])"
