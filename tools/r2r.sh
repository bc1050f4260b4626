#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 1 4 16 32 128 144 160; do
  echo "== LAV_GEMM_DEBUG=$d"
  LAV_GEMM_DEBUG=$d LAV_BENCH_GEMM_OUT=r2r_dbg$d.json timeout 200 python tools/bench_gemm.py --hot --no-cublas s1_qkv s1_fc1_gelu s1_fc2_dgrad_gelu s2_fc1_gelu s2_fc2_dgrad_gelu bert_ffn1 bert_ffn2_dgrad_gelu 2>&1 | sed -e "s/'cublas_ms': nan, 'cublas_tflops': nan//" 
done > gpurun_out/r2r_epi_ablation.log 2>&1
cat gpurun_out/r2r_epi_ablation.log
