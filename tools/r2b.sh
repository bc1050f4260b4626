#!/bin/bash
# round-2 call B: new tests (parity mode, multitask, optimizer), parity report, sanitizer follow-ups, bench, ncu DRAM
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python tools/gpu_tests.py tests > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
cp gpurun_out/pytest_all.log gpurun_out/r2b_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
timeout 600 python tools/parity_report.py gpu > gpurun_out/r2b_parity_gpu.log 2>&1
timeout 300 python tools/determinism_stress.py > gpurun_out/r2b_determinism.log 2>&1
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
# initcheck with the register epilogue (LAV_GEMM_TMA_STORE=0): are the r2a reports just untracked TMA stores?
LAV_GEMM_TMA_STORE=0 timeout 500 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_dropout_gpu.py -m gpu -q -x -p no:cacheprovider -k "encoder_train_mode" > gpurun_out/r2b_san_initcheck_notma.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_san_initcheck_notma.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider -k "bert" > gpurun_out/r2b_san_racecheck_attn.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_san_racecheck_attn.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b_san_synccheck_gemm.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_san_synccheck_gemm.log
# DRAM bytes of every GEMM launch of one step
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_f16 --csv --log-file gpurun_out/r2b_gemm_dram.csv python bench.py --profile-step > gpurun_out/r2b_ncu_dram.log 2>&1
tail -n 3 gpurun_out/r2b_tests.log; tail -n 3 gpurun_out/r2b_smoke.log; tail -n 2 gpurun_out/r2b_determinism.log; head -c 400 gpurun_out/r2b_bench.json
