#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_rowops.py gpurun_out/r2y_rowops.json 2>&1 | tee gpurun_out/r2y_rowops.log
LAV_LN_STAGED=0 timeout 300 python tools/bench_rowops.py 2>&1 | grep bert | tee gpurun_out/r2y_rowops_nostaged.log
LAV_BENCH_GEMM_OUT=r2y_gemm_head.json timeout 300 python tools/bench_gemm.py --head 2>&1 | grep "tag\|Error" | tee gpurun_out/r2y_gemm_head.log
