#!/bin/bash
# round-2 final call Z (re-run of call A on the last tree) (1 GPU): full GPU suite, smoke, the bench lines (native with all legs, reference arm, config4,
# multitask), ncu launch list + GEMM DRAM traffic of one training step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1700 python tools/gpu_tests.py tests > gpurun_out/r2Z_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2Z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2Z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2Z_bench.json 2> gpurun_out/r2Z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2Z_bench_reference.json 2> gpurun_out/r2Z_bench_reference.err; echo "ref rc=$?"
timeout 400 python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/r2Z_bench_config4.json 2> gpurun_out/r2Z_bench_config4.err
timeout 600 python bench.py --workload multitask --steps 3 > gpurun_out/r2Z_bench_multitask.json 2> gpurun_out/r2Z_bench_multitask.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2Z_launches.csv python bench.py --profile-step > gpurun_out/r2Z_ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_f16 --csv --log-file gpurun_out/r2Z_gemm_dram.csv python bench.py --profile-step > gpurun_out/r2Z_ncu_dram.log 2>&1
timeout 300 python tools/profile_step.py --out gpurun_out/r2Z_step_profile.json > gpurun_out/r2Z_step_profile.log 2>&1
tail -n 3 gpurun_out/r2Z_tests.log; tail -n 2 gpurun_out/r2Z_smoke.log
python - <<'PY'
import json
for f in ("r2Z_bench", "r2Z_bench_reference", "r2Z_bench_config4", "r2Z_bench_multitask"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("roofline", {}).get("frac"),
              (d.get("cpu_baseline") or {}).get("value"), (d.get("gpu_torch_baseline") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
wc -l gpurun_out/r2Z_launches.csv gpurun_out/r2Z_gemm_dram.csv
