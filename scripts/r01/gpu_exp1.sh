#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/chase.log
import sys; sys.path.insert(0,'bwt-merge_b200')
import bwtm_b200 as b
b.set_device(0)
for fetch in (0, 32, 64, 128):
    for table in (1<<30, 2<<30, 8<<30, 32<<30):
        for g in (32, 64, 128):
            for tps in (1024, 2048):
                r = b.chase_bench(table, g, 1<<29, tps, fetch)
                print("fetch=%3d table=%2dGB granule=%3d threads/SM=%4d : %7.1f GB/s  %6.2f G loads/s" % (fetch, table>>30, g, tps, r, r/g), flush=True)
PY
for f in 32 64 128; do
  BWTM_DEBUG=1 BWTM_L2_FETCH=$f timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -3 | tee gpurun_out/bench_fetch$f.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('FETCH $f', d['stages_ms'], d['ms_per_step'])
    else: print(l.strip())
"
done
