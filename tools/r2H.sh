#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LAV_BENCH_GEMM_OUT=r2H_gemm.json timeout 300 python tools/bench_gemm.py --hot --no-cublas s1_fc1_gelu s1_fc2_dgrad_gelu s2_fc1_gelu s2_fc2_dgrad_gelu bert_ffn1 bert_ffn2_dgrad_gelu s2_qkv 2>&1 | grep tag | cut -c1-120
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2H_tests.log 2>&1
tail -n 3 gpurun_out/r2H_tests.log
for i in 1 2; do
timeout 900 python bench.py --no-gpu-baseline --no-cpu-baseline --steps 20 > gpurun_out/r2H_bench.json 2> gpurun_out/r2H_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2H_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'], d['kernels']['gemm'])"
done
