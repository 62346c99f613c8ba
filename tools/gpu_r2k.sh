#!/bin/bash
# Round-2 eleventh GPU visit (1 GPU): TMA-staged segmented reductions -- parity first, then the A/B sweep.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_bnfold.py tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py tests/test_gpu_selftest.py tests/test_simulators.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -15 > ${P}_pytest.log
tail -n 6 ${P}_pytest.log
timeout 900 python tools/bn_sweep.py > ${P}_sweep.log 2>&1
cat ${P}_sweep.log
