#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_rowops.py gpurun_out/r2p_rowops.json > gpurun_out/r2p_rowops.log 2>&1
cat gpurun_out/r2p_rowops.log
# sanitizers over the kernels rewritten since the round-2a pass: persistent attention backward, two-thread window forward,
# deterministic grad-norm, frame transform
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py tests/test_input_pipeline.py tests/test_optim_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_san_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 4 gpurun_out/r2p_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider -k "window" > gpurun_out/r2p_san_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -n 4 gpurun_out/r2p_san_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_san_synccheck.log 2>&1
echo "synccheck rc=$?"; tail -n 4 gpurun_out/r2p_san_synccheck.log
