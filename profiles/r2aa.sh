#!/bin/bash
# r2aa: after the hull-pass fork and the acquire-RMW solver change -- multi-GPU parity on N GPUs, the world-step suite,
# per-kernel times, world-step timing on the 1M-box pile
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_single_process.py tests/test_gpu_multi.py -q -x > gpurun_out/r2aa_multi_tests.log 2>&1
echo "multi tests rc=$?"; tail -4 gpurun_out/r2aa_multi_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_world.py -q -x > gpurun_out/r2aa_world_tests.log 2>&1
echo "world tests rc=$?"; tail -3 gpurun_out/r2aa_world_tests.log | cut -c1-300
bash profiles/r2o.sh $N r2aa
timeout 300 python profiles/world_step.py --workload pile --nx 1000 --ny 1000 --steps 10 --cpu-sample 0 > gpurun_out/r2aa_world_pile.json 2> gpurun_out/r2aa_world_pile.err
python - <<PY
import json
try:
    w=json.loads([l for l in open("gpurun_out/r2aa_world_pile.json").read().splitlines() if l.startswith('{')][-1])
    print("world step pile:", w["ms_median"])
except Exception as e:
    print("world parse failed", e); print(open("gpurun_out/r2aa_world_pile.err").read()[-500:])
PY
