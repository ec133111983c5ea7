#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
BWTM_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 1 --warmup 2 2>&1 | grep -E "bwtm\[0\]|bwtm\[1\]|^\{" | tail -24 | cut -c1-400
