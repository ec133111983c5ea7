#!/bin/bash
# GPU call 2 of round 2: two-step walk + batched search: parity, sanitizer on the new kernels, bench (pairs vs single), ncu.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2.log
tail -15 gpurun_out/r02_pytest_gpu_2.log
{
echo "=== memcheck: pair records, both walks, batches"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_records_answer or both_walks or search_in_batches and 3" 2>&1 | tail -6
echo "=== racecheck: both walks"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "both_walks and not wide or pair_records_answer and reads" 2>&1 | tail -6
} > gpurun_out/r02_sanitizer_pairs.txt 2>&1
tail -12 gpurun_out/r02_sanitizer_pairs.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_pairs.json 2> gpurun_out/r02_bench_c2_pairs.err; echo "pairs rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --walk single --no-cpu-baseline > gpurun_out/r02_bench_c2_single.json 2> gpurun_out/r02_bench_c2_single.err; echo "single rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --sequence-blocks 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c2_batches4.json 2> gpurun_out/r02_bench_c2_batches4.err; echo "batches rc=$?"
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_walk_pairs|pairs_gather|pairs_fill|pairs_count' -c 5 -o gpurun_out/r02_k1_pairs python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_k1_pairs.log 2>&1; echo "ncu rc=$?"
for f in pairs single batches4; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_$f.json').read().strip().splitlines()[-1])
print('$f', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified') and d['verified']['ok'], 'walk', d.get('walk'))
PY
done
