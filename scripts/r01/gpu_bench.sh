#!/bin/bash
# Runs on the GPU box: tests, then bench at config-1 and config-2 scale.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --genome 1000000 --reads 100000 --read-len 100 --cpu-sample-reads 100000 --steps 3 --warmup 3 --verify 2>&1 | tail -5 | tee gpurun_out/bench_c1.log
timeout 1500 python bench.py --steps 3 --warmup 3 --verify --gather-bench 2>&1 | tail -5 | tee gpurun_out/bench_c2.log
