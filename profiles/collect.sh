#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root: bench line, ncu launch list of the same
# command, and one full ncu capture of the two contact kernels.  Outputs land in gpurun_out/.
# usage: profiles/collect.sh <tag> [bench args...]
set -u
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv 2>/dev/null &
SMI=$!
timeout 400 python bench.py --steps 300 --warmup 10 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step "$@" > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_manifolds|k_sweep|k_transform_aabb|k_row_map|k_bin|k_scatter_sorted|k_keys" -s 14 -c 9 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
# the device-resident world step: timing line, launch list of one step, one full capture of the solver
timeout 300 python profiles/world_step.py --workload pile --nx 1000 --ny 1000 --steps 10 --cpu-sample 100 > gpurun_out/world_$TAG.json 2> gpurun_out/world_$TAG.err
timeout 300 python profiles/world_step.py --workload polygons --nx 1000 --ny 1000 --steps 10 --cpu-sample 100 > gpurun_out/world_poly_$TAG.json 2>> gpurun_out/world_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_solve|k_chain|k_pack_rows|k_advance|k_external|DeviceRadixSort" -c 40 --csv \
    --log-file gpurun_out/launches_world_$TAG.csv python profiles/world_step.py --workload pile --nx 1000 --ny 1000 --steps 2 --warmup 1 > gpurun_out/ncu_launches_world_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_solve" -s 1 -c 1 \
    -o gpurun_out/prof_world_$TAG -f python profiles/world_step.py --workload pile --nx 1000 --ny 1000 --steps 1 --warmup 1 > gpurun_out/ncu_full_world_$TAG.log 2>&1
python profiles/_stage.py gpurun_out/bench_$TAG.json
tail -2 gpurun_out/bench_$TAG.err
# general polygons: one full capture of the 16-lanes-per-pair SAT kernel on config 2 at 1M polygons
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds_coop" -s 3 -c 1 -o gpurun_out/prof_coop_$TAG -f \
    python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/ncu_coop_$TAG.log 2>&1
