#!/bin/bash
# round-2 call C: pipelined GEMM epilogues - correctness + micro-benchmarks + step time
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c_gemm_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_gemm_tests.log
timeout 300 python tools/bench_gemm.py > gpurun_out/r2c_gemm.log 2>&1
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -n 3 gpurun_out/r2c_gemm_tests.log; tail -n 4 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_gemm.log | cut -c1-150; head -c 300 gpurun_out/r2c_bench.json
