#!/bin/bash
# round-2 call J: MMA-only rate (LAV_GEMM_DEBUG=65: no TMA loads, epilogue only releases) for the 1-CTA and the pair path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for pr in 0 1; do for d in 65 1; do
  echo "== LAV_GEMM_PAIR=$pr LAV_GEMM_DEBUG=$d" >> gpurun_out/r2j_mma_only.log
  LAV_GEMM_PAIR=$pr LAV_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py --no-cublas --sweep k512 k2048 k4096 m37888 2>&1 | cut -c1-110 >> gpurun_out/r2j_mma_only.log
  LAV_GEMM_PAIR=$pr LAV_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py --no-cublas square_8k 2>&1 | cut -c1-110 >> gpurun_out/r2j_mma_only.log
done; done
cat gpurun_out/r2j_mma_only.log
