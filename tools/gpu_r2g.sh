#!/bin/bash
# Round-2 seventh GPU visit (1 GPU): ncu evidence with the round-2 build -- launch list of one bench step, --set full
# captures of the three flat kernels in step, and of every other kernel family through tools/microbench2.py.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2g
BENCH="python bench.py --steps 2 --warmup 3 --profile"
$BENCH > ${P}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $BENCH > ${P}_ncu_launches.log 2>&1
rm -f gpurun_out/*.ncu-rep
$BENCH > ${P}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qdq_flat_hot -s 120 -c 3 -o gpurun_out/prof_qdq -f $BENCH > ${P}_ncu_qdq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reduce_flat -s 120 -c 3 -o gpurun_out/prof_reduce -f $BENCH > ${P}_ncu_reduce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_flat -s 120 -c 3 -o gpurun_out/prof_bwd -f $BENCH > ${P}_ncu_bwd.log 2>&1
B2Q_MICRO_ONCE=1 python tools/microbench2.py > ${P}_micro_once_plain.log 2>&1 &&
B2Q_MICRO_ONCE=1 ncu --set full --clock-control none -k regex:'ew_kernel|rows_cta|rows_fused|bnstat|seg_|export|reduce_flat|qdq_flat|bwd_flat|wnq|qil|dorefa|pact' -c 400 -o gpurun_out/prof_micro -f python tools/microbench2.py > ${P}_ncu_micro.log 2>&1
ncu -i gpurun_out/prof_micro.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum > ${P}_micro_raw.csv 2> ${P}_micro_raw.err
ls -la gpurun_out/*.ncu-rep
[ $(stat -c %s gpurun_out/prof_micro.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/prof_micro.ncu-rep
timeout 600 python tools/microbench2.py > ${P}_microbench2.log 2>&1
tail -n 3 ${P}_ncu_micro.log; wc -l ${P}_micro_raw.csv; grep "BN batch" ${P}_microbench2.log
