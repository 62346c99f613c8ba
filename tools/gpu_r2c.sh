#!/bin/bash
# Round-2 third GPU visit (2 GPUs): multi-rank parity, peer-exchange kernel A/B (peer_mode 1 vs 0 vs NCCL), the full default
# bench line at N=2 (what the driver runs), plus the DoReFa / selftest changes on one GPU.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2c
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_selftest.py tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_host_path.py tests/test_gpu_multi_rank.py -m gpu -q 2>&1 | tail -15 > ${P}_pytest.log
timeout 200 $RUN --master-port 29511 tests/multi_gpu_check.py > ${P}_multi_check.log 2>&1; echo "multi_check rc=$?" >> ${P}_multi_check.log
B2Q_OPT_PEER_MODE=0 timeout 200 $RUN --master-port 29512 tests/multi_gpu_check.py > ${P}_multi_check_mode0.log 2>&1; echo "multi_check rc=$?" >> ${P}_multi_check_mode0.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
timeout 300 python bench.py $SHORT > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout 300 $RUN --master-port 29513 bench.py --gpus $N $SHORT > ${P}_bench_n2_mode1.json 2> ${P}_bench_n2_mode1.err
B2Q_OPT_PEER_MODE=0 timeout 300 $RUN --master-port 29514 bench.py --gpus $N $SHORT > ${P}_bench_n2_mode0.json 2> ${P}_bench_n2_mode0.err
B2Q_EXCHANGE=nccl timeout 300 $RUN --master-port 29515 bench.py --gpus $N $SHORT > ${P}_bench_n2_nccl.json 2> ${P}_bench_n2_nccl.err
timeout 300 $RUN --master-port 29516 bench.py --gpus $N $SHORT > ${P}_bench_n2_mode1_b.json 2> ${P}_bench_n2_mode1_b.err
timeout 900 $RUN --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > ${P}_bench_n2_full.json 2> ${P}_bench_n2_full.err
timeout 300 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > ${P}_bench_ref_n2.json 2> ${P}_bench_ref_n2.err
tail -n 4 ${P}_pytest.log ${P}_multi_check.log ${P}_multi_check_mode0.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("n_gpus"), round(d.get("value",0)), d.get("ms_per_step_by_mode"), d.get("parity_checked"), (d.get("e2e") or {}).get("value"), (d.get("full_model") or {}).get("images_per_sec"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
