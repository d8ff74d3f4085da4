#!/bin/bash
# r2d: launch-bound variants of k_manifolds_coop / k_transform_aabb on 1M polygons (profiles/variants.py builds them)
set -u
mkdir -p gpurun_out
for v in "$@"; do
  SHAPES_B200_LIB=$PWD/shapes_b200/lib/var_$v.so timeout 300 python bench.py --workload polygons --shapes-per-gpu 1000000 --steps 30 --warmup 5 \
      --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/r2d_$v.json 2> gpurun_out/r2d_$v.err
  echo "== $v"; python profiles/_stage.py gpurun_out/r2d_$v.json 2>/dev/null | head -2 || tail -3 gpurun_out/r2d_$v.err
done
