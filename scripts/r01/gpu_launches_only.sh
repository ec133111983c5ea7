#!/bin/bash
# ncu launch list of one merge at config 2 (profiler range = the timed step only).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-300
