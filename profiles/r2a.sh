#!/bin/bash
# r2a: first GPU pass of round 2 -- parity suite, polygon-world stage times with the hull records in cell
# order (sorted mode) against slot order, one full ncu capture of k_manifolds_coop<true>.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
for mode in sorted slot; do
  for w in "polygons 1000000" "blob 1000000" "mixed 4000000" "polygons 10000"; do
    set -- $w
    if [ $mode = slot ]; then export SHAPES_B200_NO_SORTED=1; else unset SHAPES_B200_NO_SORTED; fi
    timeout 400 python bench.py --workload $1 --shapes-per-gpu $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-world-step \
        > gpurun_out/r2a_${mode}_$1_$2.json 2> gpurun_out/r2a_${mode}_$1_$2.err
    python profiles/_stage.py gpurun_out/r2a_${mode}_$1_$2.json 2>/dev/null | head -2 || tail -3 gpurun_out/r2a_${mode}_$1_$2.err
  done
done
unset SHAPES_B200_NO_SORTED
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds_coop|k_hulls_scatter|k_transform_aabb|k_sweep|k_scatter_sorted|k_bin|k_keys|k_rows|k_row_map" -s 9 -c 9 -o gpurun_out/prof_r2a -f \
    python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/ncu_r2a.log 2>&1
tail -3 gpurun_out/ncu_r2a.log
