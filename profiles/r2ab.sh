#!/bin/bash
# r2ab: rows mode reads the inbox records in place -- multi-GPU parity on N GPUs, per-kernel times, and
# compute-sanitizer memcheck / racecheck over a 2-GPU rows-mode frame (single-process multi ctx)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -q -x > gpurun_out/r2ab_multi_tests.log 2>&1
echo "multi tests rc=$?"; tail -4 gpurun_out/r2ab_multi_tests.log | cut -c1-300
bash profiles/r2o.sh $N r2ab
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 --print-limit 30 \
      python -m pytest tests/test_gpu_multi_single_process.py -k "oracle and 2" -x -q -p no:cacheprovider > gpurun_out/sanitize_${tool}_multi2.log 2>&1
  echo "$tool multi2 rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_${tool}_multi2.log | tail -3
done
