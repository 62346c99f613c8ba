#!/bin/bash
# Round-2 fourth GPU visit (2 GPUs): whole GPU suite (slot rings, f3 kernel, comm, fx rewriter, twins), micro-benchmarks,
# e2e with the streaming host copy, 2-GPU bench with the tighter mailbox polling.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2d
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > ${P}_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1
timeout 600 python tools/microbench2.py > ${P}_microbench2.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout 200 $RUN --master-port 29511 tests/multi_gpu_check.py > ${P}_multi_check.log 2>&1; echo "multi_check rc=$?" >> ${P}_multi_check.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
timeout 300 $RUN --master-port 29513 bench.py --gpus $N $SHORT > ${P}_bench_n2.json 2> ${P}_bench_n2.err
timeout 300 $RUN --master-port 29514 bench.py --gpus $N $SHORT > ${P}_bench_n2_b.json 2> ${P}_bench_n2_b.err
B2Q_DEBUG_SKIP_GRAD_ALLREDUCE=1 timeout 300 $RUN --master-port 29515 bench.py --gpus $N $SHORT > ${P}_bench_n2_nograd.json 2> ${P}_bench_n2_nograd.err
tail -n 12 ${P}_pytest_all.log; tail -n 2 ${P}_smoke.log ${P}_multi_check.log
grep -E "DoReFa|int8 export|V2 weight 512|fold-BN weight 512" ${P}_microbench2.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("n_gpus"), round(d.get("value",0)), d.get("ms_per_step_by_mode"), d.get("parity_checked"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("per_rank_gbs_each_direction"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
