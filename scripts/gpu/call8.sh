#!/bin/bash
# GPU call 8 of round 2 (2 GPUs): resident CTAs of the walk (A/B on one GPU), then the distributed merge with shipped planes.
mkdir -p gpurun_out
for ctas in 5 6 7; do
BWTM_WALK_CTAS=$ctas timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c2_ctas$ctas.json 2> gpurun_out/r02_bench_c2_ctas$ctas.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_ctas$ctas.json').read().strip().splitlines()[-1])
print('ctas $ctas', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'verified', d['verified']['ok'], 'warmup', d.get('warmup_ms'))
PY
done
BWTM_FINE_HISTOGRAM=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c2_fine6.json 2> gpurun_out/r02_bench_c2_fine6.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_fine6.json').read().strip().splitlines()[-1])
print('fine, 6 ctas', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()})
PY
bash scripts/gpu/dist_call.sh 2
