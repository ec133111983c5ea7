#!/bin/bash
# GPU call 3 of round 2: MSD partition sort, faster pair builder, K4 without per-key atomics.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_3.log
tail -15 gpurun_out/r02_pytest_gpu_3.log
{
echo "=== memcheck: MSD partition (config 1, two levels), counting sort, pair builder"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -k "two_msd and not True or counting_sort and 64 or pair_records_answer and reads or merge_bit_exact and reads" 2>&1 | tail -6
echo "=== racecheck: counting sort + MSD (small), K4"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "counting_sort and 1000 and not True or merge_bit_exact and repeats" 2>&1 | tail -6
} > gpurun_out/r02_sanitizer_msd.txt 2>&1
tail -12 gpurun_out/r02_sanitizer_msd.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_msd.json 2> gpurun_out/r02_bench_c2_msd.err; echo "msd rc=$?"
BWTM_MSD=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c2_cubsort.json 2> gpurun_out/r02_bench_c2_cubsort.err; echo "cub rc=$?"
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
BWTM_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k1_walk_pairs|msd_scatter|msd_histogram|k4_interleave|local_counting' -c 8 -o gpurun_out/r02_step_kernels python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pairs_gather|pairs_fill|pairs_super' -c 3 -o gpurun_out/r02_pair_builder python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
for f in msd cubsort; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c2_$f.json').read().strip().splitlines()[-1])
print('$f', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified') and d['verified']['ok'], 'pairbuild', d['walk']['pair_records_build_ms'])
PY
done
