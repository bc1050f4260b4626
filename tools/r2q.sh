#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_rowops.py gpurun_out/r2q_rowops.json > gpurun_out/r2q_rowops.log 2>&1
cat gpurun_out/r2q_rowops.log
for bn in 0 64 128 192 256; do
  LAV_GEMM_BN=$bn LAV_BENCH_GEMM_OUT=r2q_gemm_bn$bn.json timeout 300 python tools/bench_gemm.py --hot $( [ $bn != 0 ] && echo --no-cublas ) > gpurun_out/r2q_gemm_bn$bn.log 2>&1
done
python - <<'PY'
import json
t = {bn: {r["tag"]: r for r in json.load(open(f"gpurun_out/r2q_gemm_bn{bn}.json"))} for bn in (0, 64, 128, 192, 256)}
for tag, r in t[0].items():
    print(f'{tag:22s} model {r["ms"]*1e3:6.1f} us  cublas {r["cublas_ms"]*1e3:6.1f} |', "  ".join(f'bn{bn} {t[bn][tag]["ms"]*1e3:6.1f}' for bn in (64, 128, 192, 256)))
PY
