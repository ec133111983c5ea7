#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | tee gpurun_out/bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('STAGES', {k: round(v,2) for k,v in d['stages_ms'].items()}, 'ms/step', round(d['ms_per_step'],2), 'value %.3e'%d['value'], 'e2e', d['e2e'])"
