#!/bin/bash
# GPU visit (4 GPUs): 4-rank parity and the short bench of the final build (the scaling table's N=4 point).
set -u
mkdir -p gpurun_out
P=gpurun_out/r3c
N=4
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29611 tests/multi_gpu_check.py > ${P}_multi_check_4.log 2>&1; echo "rc=$?" >> ${P}_multi_check_4.log
tail -n 2 ${P}_multi_check_4.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-graph"
timeout 400 $RUN --master-port 29612 bench.py --gpus $N $SHORT > ${P}_bench_n4.json 2> ${P}_bench_n4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3c_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), (d.get("full_model") or {}).get("images_per_sec"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
