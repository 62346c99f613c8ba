#!/bin/bash
# Round-2 GPU visit (2 GPUs): peer-memory allreduce (gradients, grouped statistics) -- parity, then the A/B against NCCL.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2m
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_bnfold.py -m gpu -q -x 2>&1 | tail -12 > ${P}_pytest.log
tail -n 6 ${P}_pytest.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
B2Q_GRAD_EXCHANGE=nccl timeout 300 $RUN --master-port 29541 bench.py --gpus $N $SHORT > ${P}_n2_grad_nccl.json 2> ${P}_n2_grad_nccl.err
B2Q_GRAD_EXCHANGE=peer timeout 300 $RUN --master-port 29542 bench.py --gpus $N $SHORT > ${P}_n2_grad_peer.json 2> ${P}_n2_grad_peer.err
B2Q_GRAD_EXCHANGE=peer B2Q_OPT_PEER_ALLREDUCE_BLOCKS_PER_SM=1 timeout 300 $RUN --master-port 29543 bench.py --gpus $N $SHORT > ${P}_n2_grad_peer_b1.json 2> ${P}_n2_grad_peer_b1.err
B2Q_GRAD_EXCHANGE=nccl timeout 300 $RUN --master-port 29544 bench.py --gpus $N $SHORT > ${P}_n2_grad_nccl_b.json 2> ${P}_n2_grad_nccl_b.err
B2Q_GRAD_EXCHANGE=peer timeout 300 $RUN --master-port 29545 bench.py --gpus $N $SHORT --workload mobilenet_v1_gdrq > ${P}_n2_mobilenet_peer.json 2> ${P}_n2_mobilenet_peer.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d=json.loads(open(f).read())
        k=d.get("kernels",{})
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gradient_exchange"), {a[:10]: round(v.get("ms_total",0),2) for a,v in k.items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
