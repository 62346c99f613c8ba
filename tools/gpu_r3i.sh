#!/bin/bash
# Last 8-GPU visit of the round: 8-rank parity and the driver's N=8 command with default flags on the final tree.
set -u
mkdir -p gpurun_out
P=gpurun_out/r3i
N=8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29621 tests/multi_gpu_check.py > ${P}_multi_check_8.log 2>&1; echo "rc=$?" >> ${P}_multi_check_8.log
tail -n 2 ${P}_multi_check_8.log
( time timeout 900 $RUN --master-port 29622 bench.py --gpus $N ) > ${P}_bench_n8_full.json 2> ${P}_bench_n8_full.err
tail -n 4 ${P}_bench_n8_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3i_bench_n8_full.json").read())
print(d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gradient_exchange","")[:12], (d.get("e2e") or {}).get("value"), (d.get("full_model") or {}).get("images_per_sec"))
for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("images_per_sec"), v.get("hbm_frac_whole_step"))
PY
