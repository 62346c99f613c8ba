#!/bin/bash
set -u
mkdir -p gpurun_out
P=gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > ${P}_pytest_all.log
tail -n 3 ${P}_pytest_all.log
SHORT="--steps 20 --warmup 5 --no-e2e --no-cpu --no-micro --no-full-model"
B2Q_OPT_CLUSTER_FWD=0 timeout 600 python bench.py $SHORT > ${P}_bench_cluster0.json 2> ${P}_bench_cluster0.err
timeout 600 python bench.py $SHORT > ${P}_bench_cluster1.json 2> ${P}_bench_cluster1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, round(d.get("value",0)), {a[:14]:round(b,3) for a,b in d.get("ms_per_step_by_mode",{}).items()}, d.get("parity_checked"), d.get("gpu_launches"))
        for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("images_per_sec"), v.get("hbm_frac_whole_step"), v.get("gpu_launches_per_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
