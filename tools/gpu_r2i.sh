#!/bin/bash
# Round-2 ninth GPU visit (1 GPU): second A/B sweep of the batch-statistics kernel + one --set full capture of it and of
# the flat sum reduction (pipe utilisation: which pipe bounds the fp64 accumulation?).
set -u
mkdir -p gpurun_out
P=gpurun_out/r2i
timeout 600 python tools/bn_sweep.py > ${P}_bn_sweep.log 2>&1
cat ${P}_bn_sweep.log
B2Q_MICRO_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:'bnstat|reduce_flat_kernel<0' -c 12 -o gpurun_out/prof_bnstat -f python tools/microbench2.py > ${P}_ncu_bnstat.log 2>&1
ncu -i gpurun_out/prof_bnstat.ncu-rep --page raw --csv > ${P}_bnstat_raw.csv 2> ${P}_bnstat_raw.err
ncu -i gpurun_out/prof_bnstat.ncu-rep --page details --csv > ${P}_bnstat_details.csv 2>> ${P}_bnstat_raw.err
ls -la gpurun_out/prof_bnstat.ncu-rep
