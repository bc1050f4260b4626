#!/bin/bash
# round-2 call E: 4-stage operand ring (32 KB epilogue staging), pipelined TMA epilogue, restored register epilogue
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2e_gemm_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_gemm_tests.log
timeout 300 python tools/bench_gemm.py > gpurun_out/r2e_gemm.log 2>&1
for d in 0 1 3; do
  echo "== LAV_GEMM_DEBUG=$d" >> gpurun_out/r2e_ablation.log
  LAV_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py --no-cublas --sweep 2>&1 | cut -c1-110 >> gpurun_out/r2e_ablation.log
done
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -n 3 gpurun_out/r2e_gemm_tests.log; tail -n 4 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_gemm.log | cut -c1-150; cat gpurun_out/r2e_ablation.log; head -c 300 gpurun_out/r2e_bench.json
