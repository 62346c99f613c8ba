#!/bin/bash
# Round-2 GPU visit (8 GPUs): multi-rank parity at 8 ranks (thresholds, grouped statistics, gradient allreduce), gradient
# exchange A/B, the driver's N=8 command with default flags.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2q
N=8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29571 tests/multi_gpu_check.py > ${P}_multi_check_8.log 2>&1; echo "multi_check rc=$?" >> ${P}_multi_check_8.log
tail -n 3 ${P}_multi_check_8.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model --no-graph"
B2Q_GRAD_EXCHANGE=peer timeout 300 $RUN --master-port 29572 bench.py --gpus $N $SHORT > ${P}_n8_grad_peer.json 2> ${P}_n8_grad_peer.err
B2Q_GRAD_EXCHANGE=nccl timeout 300 $RUN --master-port 29573 bench.py --gpus $N $SHORT > ${P}_n8_grad_nccl.json 2> ${P}_n8_grad_nccl.err
( time timeout 900 $RUN --master-port 29574 bench.py --gpus $N ) > ${P}_bench_n8_full.json 2> ${P}_bench_n8_full.err
nvidia-smi topo -m > ${P}_topo.log 2>&1; lscpu | head -20 >> ${P}_topo.log
tail -n 4 ${P}_bench_n8_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gradient_exchange","")[:12], (d.get("e2e") or {}), (d.get("full_model") or {}).get("images_per_sec"))
        for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("images_per_sec"), v.get("hbm_frac_whole_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
