#!/bin/bash
# r2b: coop kernel with 8 lanes per pair, k_hulls_scatter without dependent-load chains, the new bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_circles.py -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
for w in "polygons 1000000" "blob 1000000" "mixed 4000000"; do
  set -- $w
  timeout 400 python bench.py --workload $1 --shapes-per-gpu $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-world-step \
      > gpurun_out/r2b_$1_$2.json 2> gpurun_out/r2b_$1_$2.err
  python profiles/_stage.py gpurun_out/r2b_$1_$2.json 2>/dev/null | head -3 || tail -5 gpurun_out/r2b_$1_$2.err
done
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; python profiles/_stage.py gpurun_out/r2b_bench.json || tail -5 gpurun_out/r2b_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_manifolds_coop|k_hulls_scatter|k_sweep" -s 4 -c 4 -o gpurun_out/prof_r2b -f \
    python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/ncu_r2b.log 2>&1
tail -2 gpurun_out/ncu_r2b.log
