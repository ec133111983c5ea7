#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BWTM_PROFILE_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:DeviceRadixSortOnesweepKernel -s 2 -c 2 \
   -o gpurun_out/onesweep python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_sort.log 2>&1
tail -2 gpurun_out/ncu_sort.log | cut -c1-200
