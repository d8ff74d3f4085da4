#!/bin/bash
# r2ac: final multi-GPU bench line on N ranks (full line: e2e, weak config 3, one-GPU same world), per-kernel times, and on
# 2 GPUs compute-sanitizer memcheck over the two-PROCESS rows-mode parity test
set -u
N=${1:-2}
mkdir -p gpurun_out
SHAPES_B200_KERNEL_TIMES=1 bash profiles/r2h.sh $N r2ac
grep "rank 0\] rows-mode" gpurun_out/r2ac_n$N.err | head -2 | cut -c1-1500
if [ "$N" = "2" ]; then
  timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 86 --print-limit 30 \
      python -m pytest tests/test_gpu_multi.py -k "rows and polygons" -x -q -p no:cacheprovider > gpurun_out/sanitize_memcheck_multi2proc.log 2>&1
  echo "memcheck two-process rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_multi2proc.log | tail -4
fi
