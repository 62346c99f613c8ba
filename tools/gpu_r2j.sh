#!/bin/bash
# Round-2 tenth GPU visit (1 GPU): float -> double conversions on the integer pipe: whole GPU suite, micro-benchmarks, sweep.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15 > ${P}_pytest_all.log
tail -n 5 ${P}_pytest_all.log
timeout 600 python tools/bn_sweep.py > ${P}_bn_sweep.log 2>&1
cat ${P}_bn_sweep.log
timeout 600 python tools/microbench2.py > ${P}_microbench2.log 2>&1
grep -E "meanabs|GDRQ act fwd|fold-BN data|BN batch|PACT bwd|QIL bwd|DoReFa bwd|WNQ bwd" ${P}_microbench2.log
