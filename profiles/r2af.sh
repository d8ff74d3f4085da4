#!/bin/bash
# r2af: BASELINE config 5 (1M-polygon Gaussian blob, peak density 4) on N ranks: per-rank pairs and SAT times (load balance)
set -u
N=${1:-8}
PER=$((1000000 / N))
mkdir -p gpurun_out
SHAPES_B200_KERNEL_TIMES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N \
   --workload blob --shapes-per-gpu $PER --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-world-step --no-configs > gpurun_out/r2af_n$N.json 2> gpurun_out/r2af_n$N.err
echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2af_n$N.json").read().splitlines() if l.startswith('{')][-1])
sat=d["per_rank_sat_ms"]; pr=d["per_rank_pairs"]
print("N=%d %s ms/step %.3f value %.3g" % (d["n_gpus"], d["config"]["workload"][:60], d["ms_per_step"], d["value"]))
print("  per_rank_pairs", pr, "max/mean %.3f" % (max(pr)/(sum(pr)/len(pr))))
print("  sat_ms", [round(x,3) for x in sat], "max/mean %.3f" % (max(sat)/(sum(sat)/len(sat))))
print("  stages", {k: round(v,3) for k,v in d["stage_ms"].items()})
PY
tail -3 gpurun_out/r2af_n$N.err | cut -c1-300
