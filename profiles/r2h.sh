#!/bin/bash
# r2h: multi-GPU bench lines (strong config 4 headline; weak config 3 rides along); usage: profiles/r2h.sh <ngpus> [tag]
set -u
N=${1:-2}; TAG=${2:-r2h}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 \
   > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_n$N.json").read().splitlines() if l.startswith('{')][-1])
    print("N=%d %s  ms/step %.3f  value %.3g  e2e %s" % (d["n_gpus"], d["scaling"], d["ms_per_step"], d["value"], d["e2e"] and round(d["e2e"]["ms_per_step"],2)))
    print("  stages", {k: round(v,3) for k,v in d["stage_ms"].items()})
    print("  per_rank_pairs", d["per_rank_pairs"], "sat_ms", [round(x,3) for x in d["per_rank_sat_ms"]])
    o=d.get("one_gpu_same_world"); 
    if o: print("  one gpu same world: ms %.3f" % o["ms_per_step"], {k: round(v,3) for k,v in o["stage_ms"].items()})
    w=d.get("weak_config3")
    if w: print("  weak config3: ms %.3f value %.3g" % (w.get("ms_per_step",-1), w.get("value",-1)), {k: round(v,3) for k,v in w.get("stage_ms",{}).items()}, w.get("error"))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/${TAG}_n$N.err
