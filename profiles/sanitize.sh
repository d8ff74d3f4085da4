#!/bin/bash
# compute-sanitizer passes over small GPU tests: memcheck, racecheck (shared memory hazards),
# synccheck.  Summaries land in gpurun_out/sanitize_*.log; copy the tails into profiles/.
set -u
mkdir -p gpurun_out
SMALL_PARITY='kat or stacks_variants or deleted_slots or many_vertices or sorted_mode_big or sorted_mode_crowded or empty_and_single or host_supplied or nonfinite or compact_wire or out_of_range'
SMALL_WORLD='deleted_slots or tiny_worlds or capacity or needs_an_uploaded or hot_path_call'
run() {  # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 --print-limit 30 \
      python -m pytest "$@" -x -q -p no:cacheprovider > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_${tool}_${tag}.log | tail -4
}
run memcheck parity tests/test_gpu_parity.py -k "$SMALL_PARITY"
run memcheck world tests/test_gpu_world.py tests/test_gpu_warm.py tests/test_gpu_circles.py -k "$SMALL_WORLD or cache_size or circle_kats or balls"
run racecheck parity tests/test_gpu_parity.py -k "kat or stacks_variants or sorted_mode_crowded or many_vertices"
run racecheck world tests/test_gpu_world.py -k "deleted_slots or tiny_worlds"
run synccheck parity tests/test_gpu_parity.py -k "kat or stacks_variants or sorted_mode_crowded or many_vertices"
run synccheck world tests/test_gpu_world.py -k "deleted_slots or tiny_worlds"
