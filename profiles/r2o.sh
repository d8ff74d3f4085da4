#!/bin/bash
# r2o: per-kernel CUDA-event times of rows-mode frames on N ranks (SHAPES_B200_KERNEL_TIMES=1); usage: r2o.sh <ngpus> [tag]
set -u
N=${1:-2}; TAG=${2:-r2o}
mkdir -p gpurun_out
SHAPES_B200_KERNEL_TIMES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 \
   --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$?"
grep "rows-mode kernel ms" gpurun_out/${TAG}_n$N.err | cut -c1-1800
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_n$N.json").read().splitlines() if l.startswith('{')][-1])
    print("N=%d %s  ms/step %.3f  value %.3g" % (d["n_gpus"], d["scaling"], d["ms_per_step"], d["value"]))
    print("  stages", {k: round(v,3) for k,v in d["stage_ms"].items()})
    print("  per_rank_pairs", d["per_rank_pairs"], "sat_ms", [round(x,3) for x in d["per_rank_sat_ms"]])
    o=d.get("one_gpu_same_world")
    if o: print("  one gpu same world: ms %.3f" % o["ms_per_step"])
    w=d.get("weak_config3")
    if w: print("  weak config3: ms %.3f value %.3g" % (w.get("ms_per_step",-1), w.get("value",-1)), {k: round(v,3) for k,v in w.get("stage_ms",{}).items()}, w.get("error"))
except Exception as e:
    print("parse failed", e)
PY
