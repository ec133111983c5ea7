#!/bin/bash
# GPU call 5 of round 2: config 5 (2 x 20 M x 250 bp, 5 % substitutions): the reference binary on the full workload (records
# its digest), then our arm on one GPU, checked against that digest.
mkdir -p gpurun_out
timeout 2400 python bench.py --config 5 --impl reference --record-digest --digest-out gpurun_out/r02_digest_c5.json > gpurun_out/r02_bench_reference_c5.json 2> gpurun_out/r02_bench_reference_c5.err; echo "ref c5 rc=$?"
cp tests/golden/reference_merge_digests.json gpurun_out/r02_reference_merge_digests.json
head -c 1200 gpurun_out/r02_bench_reference_c5.json; echo; tail -3 gpurun_out/r02_bench_reference_c5.err
timeout 1200 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; echo "ours c5 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c5_n1.json').read().strip().splitlines()[-1])
print('c5', 'ms', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified'), 'walk', d['walk'])
PY
tail -3 gpurun_out/r02_bench_c5_n1.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 2 --sequence-blocks 6 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_c5_n1_batches6.json 2> gpurun_out/r02_bench_c5_n1_batches6.err; echo "ours c5 batches rc=$?"
tail -c 700 gpurun_out/r02_bench_c5_n1_batches6.json
