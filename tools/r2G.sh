#!/bin/bash
# round-2 final call G: sanitizers over the kernels added at the end of the round (staged LayerNorm backward with bulk copies,
# fused gradient casts, wide-row cast, batched RMW epilogue, split-K dgrad, merged MLM head)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rowops_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2G_san_memcheck_rowops_gemm.log 2>&1
echo "memcheck rowops+gemm rc=$?"; tail -n 4 gpurun_out/r2G_san_memcheck_rowops_gemm.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_rowops_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2G_san_racecheck_rowops.log 2>&1
echo "racecheck rowops rc=$?"; tail -n 4 gpurun_out/r2G_san_racecheck_rowops.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_rowops_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2G_san_synccheck_rowops.log 2>&1
echo "synccheck rowops rc=$?"; tail -n 4 gpurun_out/r2G_san_synccheck_rowops.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider -k "merged_pass or tiny" > gpurun_out/r2G_san_memcheck_model.log 2>&1
echo "memcheck model rc=$?"; tail -n 4 gpurun_out/r2G_san_memcheck_model.log
