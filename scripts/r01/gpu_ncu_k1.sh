#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BWTM_PROFILE_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k1_walk_coop -c 1 \
   -o gpurun_out/k1_walk_coop python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/ncu_k1.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
