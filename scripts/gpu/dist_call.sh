#!/bin/bash
# Multi-GPU call of round 2 (N = $1): parity of the distributed merge (incl. injected rank-local failures), config-2 and
# config-5 bench lines, host-clock phases of one distributed merge.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
timeout 900 $RUN --master-port 29533 tests/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; tail -3 gpurun_out/r02_dist_check_n$N.log
fi
timeout 900 $RUN --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err; echo "bench c2 rc=$?"; tail -c 1800 gpurun_out/r02_bench_c2_n$N.json; echo; tail -3 gpurun_out/r02_bench_c2_n$N.err
BWTM_DEBUG=1 BWTM_BENCH_NO_DIST_E2E=1 timeout 600 $RUN --master-port 29535 bench.py --gpus $N --steps 1 --warmup 2 --no-verify > gpurun_out/r02_dist_phases_c2_n$N.json 2> gpurun_out/r02_dist_phases_c2_n$N.txt; grep "bwtm\[0\]" gpurun_out/r02_dist_phases_c2_n$N.txt | tail -12
if [ "${SKIP_C5:-0}" != "1" ]; then
timeout 1500 $RUN --master-port 29536 bench.py --gpus $N --config 5 --steps 5 --warmup 2 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err; echo "bench c5 rc=$?"; tail -c 1800 gpurun_out/r02_bench_c5_n$N.json; echo; tail -3 gpurun_out/r02_bench_c5_n$N.err
fi
if [ "${RUN_C3:-0}" == "1" ]; then
timeout 1500 $RUN --master-port 29537 scripts/config3_demo.py --steps ${C3_STEPS:-2} > gpurun_out/r02_config3_n$N.txt 2> gpurun_out/r02_config3_n$N.err; echo "config3 rc=$?"; grep -v "^\*\|OMP" gpurun_out/r02_config3_n$N.txt | tail -16 | cut -c1-400; grep -v "^\*\|OMP\|Warning" gpurun_out/r02_config3_n$N.err | tail -5
fi
