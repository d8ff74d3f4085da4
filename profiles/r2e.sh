#!/bin/bash
# r2e: + plan-ahead grid (K0 keys and bins), REDUX fold in the coop kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log
for w in "polygons 1000000" "blob 1000000" "mixed 4000000"; do
  set -- $w
  timeout 400 python bench.py --workload $1 --shapes-per-gpu $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-world-step \
      > gpurun_out/r2e_$1_$2.json 2> gpurun_out/r2e_$1_$2.err
  python profiles/_stage.py gpurun_out/r2e_$1_$2.json 2>/dev/null | head -3 || tail -5 gpurun_out/r2e_$1_$2.err
done
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; python profiles/_stage.py gpurun_out/r2e_bench.json || tail -5 gpurun_out/r2e_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds_coop|k_transform_aabb|k_sweep|k_rows|k_scatter_sorted|k_bin" -s 6 -c 6 -o gpurun_out/prof_r2e -f \
    python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/ncu_r2e.log 2>&1
tail -2 gpurun_out/ncu_r2e.log
