# usage: bash profiles/micro/solve_sweep.sh "<lanes sleep blocks prefer_i claim_after defer>" ...   (pile 1M and polygons 1M per config)
for cfg in "$@"; do
  set -- $cfg
  for wl in pile polygons; do
    echo "$wl lanes=$1 sleep=$2 blocks=$3 prefer_i=${4:-1} claim_after=${5:-0} defer=${6:-1}"
    SHAPES_B200_SOLVE_LANES=$1 SHAPES_B200_SOLVE_SLEEP=$2 SHAPES_B200_SOLVE_BLOCKS_PER_SM=$3 SHAPES_B200_SOLVE_PREFER_I=${4:-1} SHAPES_B200_SOLVE_CLAIM_AFTER=${5:-0} SHAPES_B200_SOLVE_DEFER=${6:-1} timeout 120 python profiles/world_step.py --workload $wl --nx 1000 --ny 1000 --steps 6 --warmup 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_median'], d['queue_pushes'])"
  done
done
