#!/bin/bash
# Round-2 GPU visit (2 GPUs): the driver's N=2 command with default flags (both arms), gradient exchange A/B after the
# unrolled slice-owner loop.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2o
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $RUN --master-port 29551 bench.py --gpus $N ) > ${P}_bench_n2_full.json 2> ${P}_bench_n2_full.err
( time timeout 900 $RUN --master-port 29552 bench.py --gpus $N --impl reference ) > ${P}_bench_n2_ref.json 2> ${P}_bench_n2_ref.err
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
B2Q_GRAD_EXCHANGE=nccl timeout 300 $RUN --master-port 29553 bench.py --gpus $N $SHORT > ${P}_n2_grad_nccl.json 2> ${P}_n2_grad_nccl.err
B2Q_GRAD_EXCHANGE=peer timeout 300 $RUN --master-port 29554 bench.py --gpus $N $SHORT > ${P}_n2_grad_peer.json 2> ${P}_n2_grad_peer.err
B2Q_GRAD_EXCHANGE=peer B2Q_OPT_PEER_ALLREDUCE_BLOCKS_PER_SM=4 timeout 300 $RUN --master-port 29555 bench.py --gpus $N $SHORT > ${P}_n2_grad_peer_b4.json 2> ${P}_n2_grad_peer_b4.err
timeout 200 $RUN --master-port 29556 tests/multi_gpu_check.py > ${P}_multi_check.log 2>&1; echo "multi_check rc=$?" >> ${P}_multi_check.log
tail -n 5 ${P}_bench_n2_full.err ${P}_bench_n2_ref.err; tail -n 2 ${P}_multi_check.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d=json.loads(open(f).read())
        k=d.get("kernels",{})
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gradient_exchange","")[:12], (d.get("e2e") or {}), (d.get("full_model") or {}).get("images_per_sec"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
