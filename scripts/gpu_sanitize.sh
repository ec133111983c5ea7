#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
done
echo "=== memcheck on selected parity tests"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "merge_bit_exact and (repeats or noisy) or encoder or device_builder" 2>&1 | tail -8
