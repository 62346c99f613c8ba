#!/bin/bash
# Round-2 GPU visit (1 GPU): grouped activations on small feature maps -- parity of the flattened kernels, timing.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2y
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_simulators.py -m gpu -q -x 2>&1 | tail -8 > ${P}_pytest.log
tail -n 5 ${P}_pytest.log
timeout 600 python tools/smallmap_bench.py > ${P}_smallmap.log 2>&1
cat ${P}_smallmap.log
