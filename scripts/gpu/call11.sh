#!/bin/bash
# GPU call 11 (1 GPU): final checks of the head: GPU tests, the driver's bench line, launch list and ncu of the walk.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_final.json 2> gpurun_out/r02_bench_c2_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_final.json').read().strip().splitlines()[-1])
print('final', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2), 'verified', d['verified']['ok'], 'roofline', d['roofline']['frac'], d['roofline'].get('random_records'), 'cpu', d['cpu_baseline']['value'])
PY
BWTM_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_c2_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r02_ncu_launches_final.log 2>&1; echo "launch list rc=$?"
BWTM_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k1_walk_pairs' -c 1 -o gpurun_out/r02_k1_final python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_k1_final.log 2>&1; echo "ncu rc=$?"
