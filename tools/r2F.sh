#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for sc in 1024 16384 65536; do
LAV_CHECK_DP_SCALE=$sc timeout -k 10 300 $TR --master-port 29552 bench.py --gpus 2 --check-dp > gpurun_out/r2F_check_dp_n2_s$sc.json 2> gpurun_out/r2F_check_dp_n2.err
echo "check-dp scale $sc rc=$?"
cat gpurun_out/r2F_check_dp_n2_s$sc.json
done
