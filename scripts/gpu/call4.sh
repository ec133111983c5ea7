#!/bin/bash
# GPU call 4 of round 2: tests, config-2 bench line, config 4 (1/10 scale against the reference binary, then full size).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_4.log
tail -6 gpurun_out/r02_pytest_gpu_4.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_v4.json 2> gpurun_out/r02_bench_c2_v4.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_v4.json').read().strip().splitlines()[-1])
print('c2', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified') and d['verified']['ok'], 'pairbuild', d['walk']['pair_records_build_ms'])
PY
timeout 1500 python scripts/run_config4.py --scale 0.1 --reference --reference-cli-over-abi > gpurun_out/r02_config4_tenth_vs_reference.json 2> gpurun_out/r02_config4_tenth.err; echo "config4 0.1 rc=$?"
tail -c 1500 gpurun_out/r02_config4_tenth_vs_reference.json; tail -5 gpurun_out/r02_config4_tenth.err
timeout 1500 python scripts/run_config4.py --scale 1.0 > gpurun_out/r02_config4_full.json 2> gpurun_out/r02_config4_full.err; echo "config4 full rc=$?"
tail -c 1500 gpurun_out/r02_config4_full.json; tail -5 gpurun_out/r02_config4_full.err
nvidia-smi --query-gpu=memory.used --format=csv
