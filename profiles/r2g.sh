#!/bin/bash
# r2g: rows mode on 2 GPUs -- multi-process (IPC) and single-process (shapes_multi) parity, then the whole suite on GPU 0
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py -x -q > gpurun_out/r2g_single.log 2>&1
echo "single-process rc=$?"; tail -15 gpurun_out/r2g_single.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2g_multi.log 2>&1
echo "multi-process rc=$?"; tail -15 gpurun_out/r2g_multi.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_multi_single_process.py > gpurun_out/r2g_all.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/r2g_all.log | cut -c1-300
