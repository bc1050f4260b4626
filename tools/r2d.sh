#!/bin/bash
# round-2 call D: ablation of the TMA-store epilogue (LAV_GEMM_DEBUG bits: 1 drain only, 2 no MMA, 4 no TMA store, 8 no proxy
# fence, 16 no staging/store, 32 no bias math)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 1 2 3 4 8 12 16 32 48; do
  echo "== LAV_GEMM_DEBUG=$d" >> gpurun_out/r2d_ablation.log
  LAV_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py --no-cublas --sweep 2>&1 | cut -c1-110 >> gpurun_out/r2d_ablation.log
done
cat gpurun_out/r2d_ablation.log | grep -E "==|k64|k512|swin_s2_fc1_gelu|swin_s2_fc2_res" 
