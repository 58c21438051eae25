#!/bin/bash
# memcheck over the small-scene GPU tests (films of every material / texture / instancing / sphere / patch kind, ray-cast parity)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 66 --launch-timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "tiny_scene_films or sphere_scene_films or patch_scene_films or reference_bvh or sphere_predicates or texture_lookup or film_output or options_and_depth" \
  > gpurun_out/sanitize.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize.log | head -20
