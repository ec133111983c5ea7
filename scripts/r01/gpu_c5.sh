#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# config 5 shape at 1/5 scale: 2 x 4M x 250 bp reads, 5 % substitutions, 200 Mbp genome (5x coverage)
timeout 1500 python bench.py --genome 200000000 --reads 4000000 --read-len 250 --error 0.05 --steps 2 --warmup 2 --verify --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c5_scaled.log | cut -c1-1500
# config 1 (the reference's CPU-runnable case), full size, with the CPU baseline
timeout 900 python bench.py --genome 1000000 --reads 100000 --read-len 100 --cpu-sample-reads 100000 --steps 5 --warmup 3 --verify 2>&1 | tail -1 | tee gpurun_out/bench_c1.log | cut -c1-1500
