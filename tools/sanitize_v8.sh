#!/bin/bash
# memcheck over the ABI v8 GPU tests: texture trees, texture-valued material parameters, the variety scene films
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 66 --launch-timeout 600 python -m pytest tests/test_gpu_variety.py -m gpu -x -q \
  -k "variety_scene_films or composite or constant_parameter or malformed or mapped_texture" > gpurun_out/sanitize_v8.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_v8.log | head -20
