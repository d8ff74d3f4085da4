#!/bin/bash
# Strong scaling of BASELINE config 4 (4M mixed boxes/polygons, one world) on N GPUs of one box:
#   gpurun --gpus N -- bash profiles/strong.sh N [slot-order]
# Every rank owns the pairs whose larger key lies in its slot range; with slot keys in Morton order of
# position a rank's range is a compact region of the world.
N=${1:-1}; ORDER=${2:-morton}
PER=$((4000000 / N))
ARGS="--workload mixed --shapes-per-gpu $PER --slot-order $ORDER --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-world-step"
mkdir -p gpurun_out
if [ $N -eq 1 ]; then
  timeout 300 python bench.py --gpus 1 $ARGS > gpurun_out/strong_${ORDER}_n$N.json 2> gpurun_out/strong_${ORDER}_n$N.err
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N $ARGS > gpurun_out/strong_${ORDER}_n$N.json 2> gpurun_out/strong_${ORDER}_n$N.err
fi
python profiles/_stage.py gpurun_out/strong_${ORDER}_n$N.json || tail -5 gpurun_out/strong_${ORDER}_n$N.err
