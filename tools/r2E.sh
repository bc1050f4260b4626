#!/bin/bash
# round-2 final call E: ncu --set full of the WindowAttention3D kernels (stage 2, B = 8) on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 6 -c 1 -o gpurun_out/r2E_ncu_qkv python tools/bench_gemm.py --no-cublas swin_s2_qkv > gpurun_out/r2E_ncu_qkv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 6 -c 1 -o gpurun_out/r2E_ncu_proj python tools/bench_gemm.py --no-cublas swin_s2_proj_res > gpurun_out/r2E_ncu_proj.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd_kernel|attn_bwd_kernel" -s 6 -c 2 -o gpurun_out/r2E_ncu_attn python tools/bench_attn.py win_s2 > gpurun_out/r2E_ncu_attn.log 2>&1
for f in qkv proj attn; do
  ncu -i gpurun_out/r2E_ncu_$f.ncu-rep --page raw --csv > gpurun_out/r2E_ncu_${f}_raw.csv 2>/dev/null
done
rm -f gpurun_out/r2E_ncu_*.ncu-rep
python tools/window_attention_summary.py gpurun_out/r2E_ncu_qkv_raw.csv gpurun_out/r2E_ncu_attn_raw.csv gpurun_out/r2E_ncu_proj_raw.csv "tools/r2E.sh, final tree" > gpurun_out/r2E_window_attention_ncu.json
cat gpurun_out/r2E_window_attention_ncu.json
