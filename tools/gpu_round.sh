#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), then ncu evidence.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
BENCH="python bench.py --steps 2 --warmup 3 --profile"
if [ "${1:-}" = "ncu" ]; then
  $BENCH > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
  $BENCH > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:qdq_flat_hot -s 120 -c 3 -o gpurun_out/prof_qdq -f $BENCH > gpurun_out/ncu_qdq.log 2>&1
  $BENCH > gpurun_out/plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:reduce_flat -s 120 -c 3 -o gpurun_out/prof_reduce -f $BENCH > gpurun_out/ncu_reduce.log 2>&1
  $BENCH > gpurun_out/plain4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:bwd_flat -s 120 -c 3 -o gpurun_out/prof_bwd -f $BENCH > gpurun_out/ncu_bwd.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_ours.json | head -c 3000
