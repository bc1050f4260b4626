#!/bin/bash
# round-2 call A: baseline tests + bench + microbenchmarks + compute-sanitizer on the kernel unit tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python tools/bench_attn.py --dropout > gpurun_out/r2a_attn.log 2>&1
timeout 300 python tools/bench_gemm.py > gpurun_out/r2a_gemm.log 2>&1
# compute-sanitizer: memcheck over the kernel unit tests, initcheck + racecheck over the row / dropout kernels
for t in rowops optim; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_${t}_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_san_memcheck_${t}.log 2>&1
  echo "rc=$?" >> gpurun_out/r2a_san_memcheck_${t}.log
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_san_memcheck_gemm.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_san_memcheck_gemm.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_san_memcheck_attn.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_san_memcheck_attn.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_dropout_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_san_initcheck_dropout.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_san_initcheck_dropout.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_rowops_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_san_racecheck_rowops.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_san_racecheck_rowops.log
tail -3 gpurun_out/r2a_tests.log; head -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_san_*.log
