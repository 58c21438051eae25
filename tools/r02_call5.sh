#!/bin/bash
# parity after the k_generate / k_film trims, then the occupancy A/B of the lean shade kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_variety.py -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c5_pytest.log 2>&1; tail -4 gpurun_out/r02_c5_pytest.log
rm -f gpurun_out/r02_ab_mb.log
bash tools/r02_ab_mb.sh
