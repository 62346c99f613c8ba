#!/bin/bash
set -u
mkdir -p gpurun_out
P=gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_bnfold.py -m gpu -q -x 2>&1 | tail -4 > ${P}_pytest.log
tail -n 3 ${P}_pytest.log
timeout 900 python tools/bn_sweep.py > ${P}_sweep.log 2>&1
cat ${P}_sweep.log
