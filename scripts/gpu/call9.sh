#!/bin/bash
# GPU call 9 (1 GPU): where the buffers of an index come from (second pool / default pool / cudaMalloc) and the walk's speed;
# e2e with the first input's pair records built during the second upload.
mkdir -p gpurun_out
for mode in pool default malloc; do
BWTM_RESIDENT=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_resident_$mode.json 2> gpurun_out/r02_bench_c2_resident_$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_resident_$mode.json').read().strip().splitlines()[-1])
print('resident=$mode', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2), 'verified', d['verified']['ok'], 'warmup', d.get('warmup_ms'))
PY
done
BWTM_RESIDENT=default BWTM_WALK=single timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --walk single > gpurun_out/r02_bench_c2_single_default.json 2>/dev/null
BWTM_RESIDENT=pool BWTM_WALK=single timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --walk single > gpurun_out/r02_bench_c2_single_pool.json 2>/dev/null
for f in single_default single_pool; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_$f.json').read().strip().splitlines()[-1])
print('$f', 'ms', round(d['ms_per_step'],2), 'search', round(d['stages_ms']['search'],2))
PY
done
