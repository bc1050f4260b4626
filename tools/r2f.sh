#!/bin/bash
# round-2 call F: the CTA-pair (cta_group::2) path after the staging shrink (6 stages of 32 KB) + elected tmem_empty arrives
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2f_gemm_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_gemm_tests.log
for d in 0 1 3; do
  echo "== LAV_GEMM_PAIR=1 LAV_GEMM_DEBUG=$d" >> gpurun_out/r2f_pair.log
  LAV_GEMM_PAIR=1 LAV_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py --no-cublas --sweep 2>&1 | cut -c1-110 >> gpurun_out/r2f_pair.log
done
LAV_GEMM_PAIR=1 timeout 200 python tools/bench_gemm.py --no-cublas > gpurun_out/r2f_pair_shapes.log 2>&1
timeout 200 python tools/bench_gemm.py --no-cublas > gpurun_out/r2f_single_shapes.log 2>&1
tail -n 2 gpurun_out/r2f_gemm_tests.log; cat gpurun_out/r2f_pair.log; cut -c1-120 gpurun_out/r2f_pair_shapes.log; echo SINGLE; cut -c1-120 gpurun_out/r2f_single_shapes.log
# 2-threads-per-row forward attention kernels
timeout 400 python -m pytest tests/test_attention_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2f_attn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_attn_tests.log
timeout 300 python tools/bench_attn.py --dropout > gpurun_out/r2f_attn.log 2>&1
tail -n 3 gpurun_out/r2f_attn_tests.log; cat gpurun_out/r2f_attn.log
