#!/bin/bash
# bench the non-default BASELINE configs once (1 GPU); usage: profiles/configs.sh <tag>
TAG=${1:-r1}
for w in "pile 1000000" "polygons 10000" "polygons 1000000" "blob 1000000" "mixed 4000000"; do
  set -- $w
  timeout 400 python bench.py --workload $1 --shapes-per-gpu $2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/cfg_${TAG}_$1_$2.json 2> gpurun_out/cfg_${TAG}_$1_$2.err
  python profiles/_stage.py gpurun_out/cfg_${TAG}_$1_$2.json 2>/dev/null | head -2 || tail -3 gpurun_out/cfg_${TAG}_$1_$2.err
done
