#!/bin/bash
# compute-sanitizer over the smoke merge (three tools) and memcheck / racecheck over selected parity tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
for tool in memcheck racecheck initcheck; do
  echo "=== $tool: smoke()"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
done
SEL="merge_bit_exact and (repeats or noisy) or encoder or device_builder or counting_sort or result_index or create_pair or streaming"
echo "=== memcheck on selected parity tests"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" 2>&1 | tail -6
echo "=== racecheck on selected parity tests"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "merge_bit_exact and repeats or counting_sort and not wide or encoder" 2>&1 | tail -6
} 2>&1 | tee gpurun_out/sanitizer.txt
