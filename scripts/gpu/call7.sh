#!/bin/bash
# GPU call 7 of round 2 (1 GPU): two memory pools, fine histogram from the walk (A/B), config 5 in one shot, launch lists.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_7.log
tail -6 gpurun_out/r02_pytest_gpu_7.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_fine.json 2> gpurun_out/r02_bench_c2_fine.err; echo "fine rc=$?"
BWTM_FINE_HISTOGRAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c2_coarse.json 2> gpurun_out/r02_bench_c2_coarse.err; echo "coarse rc=$?"
timeout 1200 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; echo "c5 rc=$?"
for f in c2_fine c2_coarse c5_n1; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1])
print('$f', 'ms', round(d['ms_per_step'],2), 'steps', d['steps_ms'], 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified') and d['verified']['ok'], 'batches', d['walk']['search_batches'], 'mem', d['device_memory'])
PY
done
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list c2 rc=$?"
BWTM_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_c5.csv python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r02_ncu_launches_c5.log 2>&1; echo "launch list c5 rc=$?"
python scripts/launch_summary.py gpurun_out/r02_launches_c5.csv | head -14
