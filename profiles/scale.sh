#!/bin/bash
# weak-scaling sweep on one box: N = 1, 2, 4, 8 (1M shapes per GPU); run under `gpurun --gpus 8`
TAG=${1:-r1}
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  python profiles/_stage.py gpurun_out/scale_${TAG}_n$N.json || tail -5 gpurun_out/scale_${TAG}_n$N.err
done
