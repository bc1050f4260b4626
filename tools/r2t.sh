#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py --out gpurun_out/r2t_step_profile.json > gpurun_out/r2t_step_profile.log 2>&1
grep -v Warning gpurun_out/r2t_step_profile.log | tail -n 150
