#!/bin/bash
# GPU tests, short bench, then the ncu launch list of one step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash scripts/gpu_quick.sh "$@"
bash scripts/gpu_launches_only.sh > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/launches.csv | grep -v "cub::DeviceScanInit" | head -${LL_LINES:-14}
