#!/bin/bash
# r2f: final single-GPU pass of round 2 -- the whole 1-GPU suite, the default bench line (with the configs sub-records),
# configs sub-records), the ncu launch list of the same command and one full capture of the polygon-world kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; python profiles/_stage.py gpurun_out/r2f_bench.json || tail -5 gpurun_out/r2f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r2f_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 80 --csv \
    --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/r2f_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 80 --csv \
    --log-file gpurun_out/r2f_launches_polygons.csv python bench.py --workload polygons --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/r2f_ncu_launches_poly.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds|k_hulls|k_transform_aabb|k_sweep|k_rows|k_scatter_sorted|k_bin|k_row_map" -s 12 -c 12 -o gpurun_out/prof_r2f_poly -f \
    python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/r2f_ncu_poly.log 2>&1
tail -2 gpurun_out/r2f_ncu_poly.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds|k_transform_aabb|k_sweep|k_rows|k_row_map" -s 10 -c 8 -o gpurun_out/prof_r2f_pile -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/r2f_ncu_pile.log 2>&1
tail -2 gpurun_out/r2f_ncu_pile.log
