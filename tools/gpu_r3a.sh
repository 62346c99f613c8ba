#!/bin/bash
# GPU visit (2 GPUs): peer_mode 4 (the reduction's last block publishes) -- parity, then A/B against peer_mode 1.
set -u
mkdir -p gpurun_out
P=gpurun_out/r3a
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi_rank.py "tests/test_gpu_fullsize.py::test_peer_kernels_world1_bit_exact_at_full_size" -m gpu -q -x 2>&1 | tail -4 > ${P}_pytest.log
tail -n 2 ${P}_pytest.log
B2Q_OPT_PEER_MODE=4 timeout 200 $RUN --master-port 29590 tests/multi_gpu_check.py > ${P}_multi_check_mode4.log 2>&1; echo "rc=$?" >> ${P}_multi_check_mode4.log
tail -n 2 ${P}_multi_check_mode4.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
port=29591
for M in 1 4 1 4; do
  port=$((port+1))
  B2Q_OPT_PEER_MODE=$M timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode${M}_$port.json 2> ${P}_n2_mode${M}_$port.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3a_*.json")):
    try:
        d=json.loads(open(f).read())
        k=d.get("kernels",{})
        print(f, round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), {a[:10]: round(v.get("ms_total",0),2) for a,v in k.items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
