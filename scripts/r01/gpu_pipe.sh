#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "1 8" "2 6" "4 6" "4 5" "4 8" "8 6" "8 5"; do
  set -- $cfg
  BWTM_PIPELINE_CHUNKS=$1 BWTM_WALK_CTAS=$2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('CHUNKS $1 CTAS $2', {k: round(v,2) for k,v in d['stages_ms'].items()}, 'ms/step', round(d['ms_per_step'],2))"
done
