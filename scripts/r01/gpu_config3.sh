#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}; shift
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 scripts/config3_demo.py "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tail -60 | tee gpurun_out/config3_n$N.log
