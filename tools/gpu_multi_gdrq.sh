#!/bin/bash
# Multi-GPU visit for the mean-based thresholds (BASELINE.json config 4): parity script, then MobileNet-v1 GDRQ_PY with
# the NCCL exchange and with the fused peer-memory exchange.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $RUN --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi_check rc=$?" >> gpurun_out/multi_check_$N.log
tail -4 gpurun_out/multi_check_$N.log
for X in nccl peer; do
  B2Q_EXCHANGE=$X timeout 200 $RUN --master-port 29514 bench.py --gpus $N --workload mobilenet_v1_gdrq --steps 10 --warmup 3 \
      --no-cpu --no-e2e > gpurun_out/bench_${N}gpu_mobilenet_v1_gdrq_$X.json 2> gpurun_out/bench_${N}gpu_mobilenet_v1_gdrq_$X.err
  python - "gpurun_out/bench_${N}gpu_mobilenet_v1_gdrq_$X" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1] + ".json"))
    print(sys.argv[1], d["n_gpus"], round(d["value"]), d["ms_per_step"], d["config"]["threshold_exchange"], d["ms_per_step_by_mode"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open(sys.argv[1] + ".err").read()[-1500:])
PY
done
