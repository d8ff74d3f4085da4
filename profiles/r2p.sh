#!/bin/bash
# r2p: after a rows-mode change -- both multi-GPU parity suites on 2 GPUs, then the per-kernel times (r2o)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -q -x > gpurun_out/r2p_multi_tests.log 2>&1
echo "multi tests rc=$?"; tail -4 gpurun_out/r2p_multi_tests.log | cut -c1-300
bash profiles/r2o.sh ${1:-2} ${2:-r2p}
