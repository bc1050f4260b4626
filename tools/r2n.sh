#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/trace_attn.py > gpurun_out/r2n_trace_attn.log 2>&1
cat gpurun_out/r2n_trace_attn.log
