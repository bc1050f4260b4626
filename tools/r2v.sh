#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1700 python tools/gpu_tests.py tests > gpurun_out/r2v_gpu_tests.log 2>&1
tail -n 12 gpurun_out/r2v_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/r2v_smoke.log
