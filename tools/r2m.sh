#!/bin/bash
# round-2 call M: ncu --set full with source-level stall sampling of the window attention backward / forward kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_bwd_kernel" -s 3 -c 1 -o gpurun_out/r2m_bwd python tools/bench_attn.py win_s2 > gpurun_out/r2m_ncu_bwd.log 2>&1
timeout 300 python tools/ncu_stalls.py gpurun_out/r2m_bwd.ncu-rep 45 > gpurun_out/r2m_bwd_stalls.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_bwd_kernel" -s 3 -c 1 -o gpurun_out/r2m_bwd_bert python tools/bench_attn.py bert_vtm --dropout > gpurun_out/r2m_ncu_bwd_bert.log 2>&1
timeout 300 python tools/ncu_stalls.py gpurun_out/r2m_bwd_bert.ncu-rep 45 > gpurun_out/r2m_bwd_bert_stalls.txt 2>&1
rm -f gpurun_out/*.ncu-rep
head -60 gpurun_out/r2m_bwd_stalls.txt
