#!/bin/bash
# Round-end evidence: GPU tests, smoke, bench (both arms), ncu launch list over one merge, ncu full capture of K1.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-400 gpurun_out/bench_ours.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
BWTM_PROFILE_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k1_walk_coop -c 1 \
   -f -o gpurun_out/k1_walk_coop python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out
