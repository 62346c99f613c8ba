#!/bin/bash
# Multi-GPU visit for BASELINE.json configs 4 and 5 (gpurun --gpus N): the multi-rank parity script (minmax + GDRQ
# thresholds), then MobileNet-v1 GDRQ_PY and ResNeXt-101 ClipGrad data parallel, then the default 1-GPU bench line.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_config1.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_multi_rank.log
timeout 150 $RUN --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi_check rc=$?" >> gpurun_out/multi_check_$N.log
for W in mobilenet_v1_gdrq resnext101_clipgrad; do
  B=256; [ $W = resnext101_clipgrad ] && B=128
  timeout 240 $RUN --master-port 29514 bench.py --gpus $N --workload $W --batch $B --steps 10 --warmup 3 --no-cpu --no-e2e \
      > gpurun_out/bench_${N}gpu_$W.json 2> gpurun_out/bench_${N}gpu_$W.err
done
timeout 400 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/pytest_multi_rank.log; tail -3 gpurun_out/multi_check_$N.log
for W in mobilenet_v1_gdrq resnext101_clipgrad; do
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
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ours.json"))
print(round(d["value"]), d["ms_per_step"], d.get("full_model"))
PY
