#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/cluster_sweep.py > gpurun_out/r2s_cluster_sweep.log 2>&1
cat gpurun_out/r2s_cluster_sweep.log
