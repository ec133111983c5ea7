#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 1 2; do
  BWTM_WALK=$v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_walk$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('WALK $v', d['stages_ms'], 'ms/step', d['ms_per_step'], 'value', d['value'])"
done
