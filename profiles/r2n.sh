#!/bin/bash
# r2n: 2-GPU lease -- both multi-GPU parity suites (log kept), the in-place growth tests, the N=2 bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -q -rA > gpurun_out/r2n_multi_tests.log 2>&1
echo "multi tests rc=$?"; tail -25 gpurun_out/r2n_multi_tests.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -k "capacity or extent" > gpurun_out/r2n_grow_tests.log 2>&1
echo "grow tests rc=$?"; tail -15 gpurun_out/r2n_grow_tests.log | cut -c1-300
bash profiles/r2h.sh 2 r2n
