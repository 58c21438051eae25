#!/bin/bash
# racecheck (shared-memory hazards: the traversal stacks, the per-chunk counting sort of the textured shade kernels) and
# synccheck (barrier / warp-sync misuse: the ballot-aggregated queue appends, the voted phase loop) over small renders of
# every kernel family -- lean, textured, general (instances / spheres), Mix -- plus the film atomics.  SURVEY section 5.
mkdir -p gpurun_out
: > gpurun_out/r02_sanitize_race.log
for TOOL in racecheck synccheck; do
  for K in "tiny_scene_films and (diffuse or texewa or inst or coated)" "variety_scene_films and (mix or envmap)"; do
    echo "== $TOOL :: $K" >> gpurun_out/r02_sanitize_race.log
    timeout 500 compute-sanitizer --tool $TOOL --error-exitcode 66 --launch-timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variety.py -m gpu -x -q -k "$K" \
      >> gpurun_out/r02_sanitize_race.log 2>&1
    echo "exit $?" >> gpurun_out/r02_sanitize_race.log
  done
done
grep -E "^== |^exit|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|passed|failed|hazard" gpurun_out/r02_sanitize_race.log | head -60
