#!/bin/bash
# Runs on the GPU box: bench at config 2, gather microbenchmark, ncu launch list and a full capture of K1.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 --gather-bench 2>&1 | tail -3 | tee gpurun_out/bench_c2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k1_walk -s 1 -c 1 -o gpurun_out/k1_walk \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out
