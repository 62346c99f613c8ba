#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): parity of the threshold exchange, then the bench with both exchange paths.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $RUN --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi_check rc=$?" >> gpurun_out/multi_check_$N.log
B2Q_EXCHANGE=nccl timeout 200 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --e2e-steps 2 \
    > gpurun_out/bench_${N}gpu_nccl.json 2> gpurun_out/bench_${N}gpu_nccl.err
B2Q_EXCHANGE=peer timeout 200 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --e2e-steps 2 \
    > gpurun_out/bench_${N}gpu_peer.json 2> gpurun_out/bench_${N}gpu_peer.err
tail -3 gpurun_out/multi_check_$N.log
for f in gpurun_out/bench_${N}gpu_nccl gpurun_out/bench_${N}gpu_peer; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1] + ".json"))
    print(sys.argv[1], d["n_gpus"], round(d["value"]), d["ms_per_step"], d["config"]["threshold_exchange"], d["ms_per_step_by_mode"], d.get("e2e", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open(sys.argv[1] + ".err").read()[-1500:])
PY
done
