#!/bin/bash
# GPU call 1 of round 2: tests, both bench arms (records the config-2 reference digest), config 1, ncu of K4/K5.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt; df -h /dev/shm >> gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1.log
tail -5 gpurun_out/r02_pytest_gpu_1.log
timeout 600 python bench.py --impl reference --record-digest --digest-out gpurun_out/r02_digest_c2.json > gpurun_out/r02_bench_reference_c2.json 2> gpurun_out/r02_bench_reference_c2.err; echo "ref rc=$?"
cp tests/golden/reference_merge_digests.json gpurun_out/r02_reference_merge_digests.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_ours_c2_first.json 2> gpurun_out/r02_bench_ours_c2_first.err; echo "ours rc=$?"
timeout 300 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_ours_c1.json 2> gpurun_out/r02_bench_ours_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --config 1 --impl reference > gpurun_out/r02_bench_reference_c1.json 2> gpurun_out/r02_bench_reference_c1.err; echo "ref c1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k4_interleave|enc_write_tiles|enc_collect_long|run_tile_survey|k4_partition|enc_tile_maps' -c 18 -o gpurun_out/r02_k4k5 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_k4k5.log 2>&1; echo "ncu rc=$?"
head -c 1500 gpurun_out/r02_bench_ours_c2_first.json; echo; head -c 1500 gpurun_out/r02_bench_reference_c2.json; echo
