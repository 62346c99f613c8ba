#!/bin/bash
# Round-2 GPU visit (1 GPU): single-launch cluster forward for small tensors -- parity, whole suite, bench A/B.
set -u
mkdir -p gpurun_out
P=gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x 2>&1 | tail -15 > ${P}_pytest_cluster.log
tail -n 8 ${P}_pytest_cluster.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > ${P}_pytest_all.log
tail -n 3 ${P}_pytest_all.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-micro --no-full-model"
B2Q_OPT_CLUSTER_FWD=0 timeout 600 python bench.py $SHORT > ${P}_bench_cluster0.json 2> ${P}_bench_cluster0.err
B2Q_OPT_CLUSTER_FWD=1 timeout 600 python bench.py $SHORT > ${P}_bench_cluster1.json 2> ${P}_bench_cluster1.err
B2Q_OPT_CLUSTER_FWD=1 B2Q_OPT_CLUSTER_WORDS_PER_CTA=512 timeout 600 python bench.py $SHORT --no-workloads > ${P}_bench_cluster1_w512.json 2> ${P}_bench_cluster1_w512.err
B2Q_OPT_CLUSTER_FWD=1 B2Q_OPT_CLUSTER_WORDS_PER_CTA=2048 timeout 600 python bench.py $SHORT --no-workloads > ${P}_bench_cluster1_w2048.json 2> ${P}_bench_cluster1_w2048.err
timeout 600 python tools/microbench2.py 2>&1 | grep -E "weight" > ${P}_micro_weights.log
cat ${P}_micro_weights.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gpu_launches"))
        for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("images_per_sec"), v.get("hbm_frac_whole_step"), v.get("gpu_launches_per_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
