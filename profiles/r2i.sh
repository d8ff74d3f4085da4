#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -x -q > gpurun_out/r2i_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2i_tests.log | cut -c1-300
bash profiles/r2h.sh 2 r2i
