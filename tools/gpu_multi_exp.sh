#!/bin/bash
# Experiment (2 GPUs): where does the multi-GPU overhead come from?  (a) as shipped, (b) gradient allreduce skipped,
# (c) NCCL limited to a few CTAs so that it takes fewer SMs away from the backward sweeps.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 200 $RUN --master-port 29516 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e \
      > gpurun_out/exp_${N}gpu_$label.json 2> gpurun_out/exp_${N}gpu_$label.err
  python - "gpurun_out/exp_${N}gpu_$label" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1] + ".json"))
    print(sys.argv[1], round(d["value"]), d["ms_per_step"], d["ms_per_step_by_mode"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1] + ".err").read()[-800:])
PY
}
run shipped B2Q_EXCHANGE=peer
run nograd B2Q_EXCHANGE=peer B2Q_DEBUG_SKIP_GRAD_ALLREDUCE=1
run ctas4 B2Q_EXCHANGE=peer NCCL_MAX_CTAS=4
run ctas8 B2Q_EXCHANGE=peer NCCL_MAX_CTAS=8
