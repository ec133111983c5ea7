#!/bin/bash
# ncu --set full capture of one launch of the kernel whose name matches $1 (regex), inside one bench step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
K=${1:-k1_walk_coop}; OUT=${2:-$K}
BWTM_PROFILE_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -c 1 \
   -f -o gpurun_out/$OUT python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$OUT.log 2>&1
tail -2 gpurun_out/ncu_$OUT.log | cut -c1-200
ls -la gpurun_out/$OUT.ncu-rep
