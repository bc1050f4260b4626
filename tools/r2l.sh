#!/bin/bash
# round-2 call L: persistent attention backward (prefetch across work items) + full-width configs[3] golden
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_attention_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_attn_tests.log 2>&1
echo "attn tests rc=$?" | tee -a gpurun_out/r2l_attn_tests.log
timeout 300 python tools/bench_attn.py --dropout > gpurun_out/r2l_attn.log 2>&1
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -n 3 gpurun_out/r2l_attn_tests.log; cat gpurun_out/r2l_attn.log; tail -n 4 gpurun_out/r2l_tests.log; grep -h "large384\|high-precision" gpurun_out/pytest_all.log | head; head -c 250 gpurun_out/r2l_bench.json
