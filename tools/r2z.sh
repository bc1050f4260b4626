#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LAV_BENCH_GEMM_OUT=r2z_gemm_head.json timeout 300 python tools/bench_gemm.py --head --no-cublas dec_wgrad_128 dec_wgrad_32 dec_wgrad_160 2>&1 | grep "tag\|Error"
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1
tail -n 3 gpurun_out/r2z_tests.log
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])"
