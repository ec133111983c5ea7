#!/bin/bash
# Distributed parity (three exchange routes) + phase clocks + bench at N GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12 | tee gpurun_out/dist_check.log
BWTM_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 1 --warmup 3 2>&1 | grep -E "bwtm\[0\]|bwtm\[1\]|Error|error" | tail -26 | cut -c1-300 | tee gpurun_out/dist_phases_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee gpurun_out/bench_n$N.log | cut -c1-1200
