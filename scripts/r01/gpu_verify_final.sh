#!/bin/bash
# Final full-size checks: config 2 with --verify (merged bytes == direct construction of BWT(A ++ B)) as the bench line of
# record, and the config-5 shape at 1/5 scale with --verify.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --verify 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-300 gpurun_out/bench_ours.json; grep -o '"verify": "[^"]*"' gpurun_out/bench_ours.json
timeout 900 python bench.py --genome 200000000 --reads 4000000 --read-len 250 --error 0.05 --steps 2 --warmup 2 --verify --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5_scaled.log; cut -c1-300 gpurun_out/bench_c5_scaled.log; grep -o '"verify": "[^"]*"\|"stages_ms": {[^}]*}' gpurun_out/bench_c5_scaled.log
