#!/bin/bash
# round-2 call G: full tests, bench, config4 / multitask lines, ncu --set full of the WindowAttention3D kernels (stage 2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 400 python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/r2g_bench_config4.json 2> gpurun_out/r2g_bench_config4.err
timeout 600 python bench.py --workload multitask --steps 3 > gpurun_out/r2g_bench_multitask.json 2> gpurun_out/r2g_bench_multitask.err
# ncu --set full: qkv GEMM, window attention fwd/bwd, proj GEMM at Swin stage 2, B = 8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 6 -c 1 -o gpurun_out/r2g_ncu_qkv python tools/bench_gemm.py --no-cublas swin_s2_qkv > gpurun_out/r2g_ncu_qkv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 6 -c 1 -o gpurun_out/r2g_ncu_proj python tools/bench_gemm.py --no-cublas swin_s2_proj_res > gpurun_out/r2g_ncu_proj.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd_kernel|attn_bwd_kernel" -s 6 -c 2 -o gpurun_out/r2g_ncu_attn python tools/bench_attn.py win_s2 > gpurun_out/r2g_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_f16 -s 6 -c 1 -o gpurun_out/r2g_ncu_ffn1 python tools/bench_gemm.py --no-cublas bert_ffn1_vtm > gpurun_out/r2g_ncu_ffn1.log 2>&1
for f in qkv proj attn ffn1; do
  ncu -i gpurun_out/r2g_ncu_$f.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_${f}_raw.csv 2>/dev/null
done
rm -f gpurun_out/r2g_ncu_*.ncu-rep
tail -n 4 gpurun_out/r2g_tests.log; head -c 300 gpurun_out/r2g_bench.json; echo; head -c 600 gpurun_out/r2g_bench_config4.json; echo; cat gpurun_out/r2g_bench_multitask.json; tail -n 3 gpurun_out/r2g_bench_multitask.err; ls -la gpurun_out/r2g_ncu_*
