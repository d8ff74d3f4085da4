for w in "polygons 1000000 generator" "blob 1000000 generator" "mixed 4000000 generator" "mixed 4000000 morton" "polygons 1000000 morton"; do
  set -- $w
  timeout 400 python bench.py --workload $1 --shapes-per-gpu $2 --slot-order $3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-world-step > gpurun_out/cfg_h_$1_$2_$3.json 2> gpurun_out/cfg_h_$1_$2_$3.err
  echo "$1 $2 $3: $(python profiles/_stage.py gpurun_out/cfg_h_$1_$2_$3.json 2>/dev/null | grep -o 'ms/step [0-9.]*\|transform_aabb [0-9.]*\|manifolds [0-9.]*\|contact_rows [0-9.]*' | tr '\n' ' ')"
done
