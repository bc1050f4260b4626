#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_rowops_gpu.py -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1
tail -n 4 gpurun_out/r2w_tests.log
LAV_BENCH_GEMM_OUT=r2w_gemm.json timeout 300 python tools/bench_gemm.py --hot --no-cublas s1_fc2_dgrad_gelu s2_fc2_dgrad_gelu bert_ffn2_dgrad_gelu s2_fc1_gelu 2>&1 | grep tag
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_dropout_gpu.py -m gpu -x -q > gpurun_out/r2w_model_tests.log 2>&1
tail -n 3 gpurun_out/r2w_model_tests.log
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2w_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items()})"
