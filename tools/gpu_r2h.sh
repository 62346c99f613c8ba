#!/bin/bash
# Round-2 eighth GPU visit (1 GPU): batch-statistics + fold kernel after the register diet -- parity, then the A/B sweep.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2h
timeout 600 python -m pytest tests/test_gpu_bnfold.py tests/test_gpu_host_path.py -m gpu -q -x 2>&1 | tail -8 > ${P}_pytest.log
tail -n 4 ${P}_pytest.log
timeout 600 python tools/bn_sweep.py > ${P}_bn_sweep.log 2>&1
cat ${P}_bn_sweep.log
