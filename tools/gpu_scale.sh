#!/bin/bash
# Weak-scaling point at N GPUs (charged N x): parity script + the ResNet-50 bench line (no e2e / cpu legs).
N=${1:-4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 120 $RUN --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi_check rc=$?" >> gpurun_out/multi_check_$N.log
tail -2 gpurun_out/multi_check_$N.log
for W in ${2:-resnet50_int8}; do
  timeout 150 $RUN --master-port 29514 bench.py --gpus $N --workload $W --steps 20 --warmup 3 --no-cpu --no-e2e \
      > gpurun_out/bench_${N}gpu_$W.json 2> gpurun_out/bench_${N}gpu_$W.err
  python - "gpurun_out/bench_${N}gpu_$W" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1] + ".json"))
    print(sys.argv[1], d["n_gpus"], round(d["value"]), d["ms_per_step"], d["config"]["threshold_exchange"], d["ms_per_step_by_mode"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open(sys.argv[1] + ".err").read()[-1500:])
PY
done
