#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -x -q > gpurun_out/r2l_tests.log 2>&1
echo "multi tests rc=$?"; tail -6 gpurun_out/r2l_tests.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_multi_single_process.py > gpurun_out/r2l_all.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/r2l_all.log | cut -c1-300
bash profiles/r2h.sh 2 r2l
