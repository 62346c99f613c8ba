#!/bin/bash
set -u
mkdir -p gpurun_out
P=gpurun_out/r3b
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model --no-graph"
port=29600
for B in 4 8 16 32; do
  port=$((port+1))
  B2Q_OPT_PEER_MODE=4 B2Q_OPT_PEER_PUBLISH_BLOCKS_PER_SM=$B timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode4_b$B.json 2> ${P}_n2_mode4_b$B.err
done
port=$((port+1))
B2Q_OPT_PEER_MODE=1 timeout 300 $RUN --master-port $port bench.py --gpus $N $SHORT > ${P}_n2_mode1.json 2> ${P}_n2_mode1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3b_*.json")):
    try:
        d=json.loads(open(f).read())
        k=d.get("kernels",{})
        print(f, round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), {a[:10]: round(v.get("ms_total",0),2) for a,v in k.items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
