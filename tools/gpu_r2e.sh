#!/bin/bash
# Round-2 fifth GPU visit (1 GPU): whole GPU suite again after the r2d fix, host-copy probe for the e2e leg.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > ${P}_pytest_all.log
timeout 600 tools/hostcopy_probe 2048 > ${P}_hostcopy_probe.log 2>&1
lscpu > ${P}_lscpu.log 2>&1; numactl -H >> ${P}_lscpu.log 2>&1; nvidia-smi topo -m >> ${P}_lscpu.log 2>&1
tail -n 6 ${P}_pytest_all.log; cat ${P}_hostcopy_probe.log
