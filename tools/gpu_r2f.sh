#!/bin/bash
# Round-2 sixth GPU visit (2 GPUs): staged peer sweep (peer_mode 2 / 3) vs peer_mode 1, at world 1 and at 2 GPUs; e2e with
# 512-bit streaming stores.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2f
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_host_path.py "tests/test_gpu_fullsize.py::test_peer_kernels_world1_bit_exact_at_full_size" -m gpu -q -x --durations=5 2>&1 | tail -25 > ${P}_pytest.log
tail -n 8 ${P}_pytest.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
timeout 300 python bench.py $SHORT > ${P}_n1_plain.json 2> ${P}_n1_plain.err
for M in 1 2 3; do
  B2Q_DEBUG_PEER_WORLD1=1 B2Q_OPT_PEER_MODE=$M timeout 300 python bench.py $SHORT > ${P}_w1_mode$M.json 2> ${P}_w1_mode$M.err
done
B2Q_DEBUG_PEER_WORLD1=1 B2Q_OPT_PEER_MODE=2 B2Q_OPT_PEER_STAGE_EARLY=1 timeout 300 python bench.py $SHORT > ${P}_w1_mode2e.json 2> ${P}_w1_mode2e.err
port=29520
for M in 1 2 3; do
  port=$((port+1))
  B2Q_OPT_PEER_MODE=$M timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode$M.json 2> ${P}_n2_mode$M.err
  port=$((port+1))
  B2Q_OPT_PEER_MODE=$M timeout 200 $RUN --master-port $port tests/multi_gpu_check.py > ${P}_multi_check_mode$M.log 2>&1; echo "multi_check mode $M rc=$?" >> ${P}_multi_check_mode$M.log
done
port=$((port+1))
B2Q_OPT_PEER_MODE=2 B2Q_OPT_PEER_STAGE_EARLY=1 timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode2e.json 2> ${P}_n2_mode2e.err
port=$((port+1))
B2Q_OPT_PEER_MODE=1 timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode1_b.json 2> ${P}_n2_mode1_b.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-workloads --no-micro --no-full-model > ${P}_n1_e2e.json 2> ${P}_n1_e2e.err
tail -n 2 ${P}_multi_check_mode*.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        d=json.loads(open(f).read())
        k=d.get("kernels",{})
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), (d.get("e2e") or {}).get("value"), {a[:10]: round(v.get("ms_total",0),2) for a,v in k.items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
