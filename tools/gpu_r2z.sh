#!/bin/bash
# Round-2 last 2-GPU visit: 2-rank parity and the driver's N=2 command with default flags on the final build.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2z
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_comm.py -m gpu -q -x 2>&1 | tail -4 > ${P}_pytest.log
tail -n 2 ${P}_pytest.log
( time timeout 900 $RUN --master-port 29581 bench.py --gpus $N ) > ${P}_bench_n2_full.json 2> ${P}_bench_n2_full.err
timeout 300 $RUN --master-port 29582 bench.py --gpus $N --workload mobilenet_v1_gdrq --steps 20 --warmup 5 --no-e2e --no-cpu --no-micro --no-full-model > ${P}_bench_n2_mobilenet.json 2> ${P}_bench_n2_mobilenet.err
timeout 300 $RUN --master-port 29583 bench.py --gpus $N --workload resnext101_clipgrad --batch 128 --steps 10 --warmup 3 --no-e2e --no-cpu --no-micro --no-full-model > ${P}_bench_n2_resnext.json 2> ${P}_bench_n2_resnext.err
tail -n 4 ${P}_bench_n2_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("n_gpus"), round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gradient_exchange","")[:12], (d.get("e2e") or {}).get("value"), (d.get("full_model") or {}).get("images_per_sec"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
