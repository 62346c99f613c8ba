#!/bin/bash
# Round-2 GPU visit (1 GPU): full evidence pass of the current build -- whole GPU suite, smoke, both bench arms with default
# flags (timed with `time` to check the driver's budget), micro-benchmarks.
set -u
mkdir -p gpurun_out
P=gpurun_out/r3f
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -15 > ${P}_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1
( time timeout 900 python bench.py --impl reference ) > ${P}_bench_reference.json 2> ${P}_bench_reference.err
( time timeout 900 python bench.py ) > ${P}_bench_ours.json 2> ${P}_bench_ours.err
timeout 600 python tools/microbench2.py > ${P}_microbench2.log 2>&1
tail -n 4 ${P}_pytest_all.log; tail -n 2 ${P}_smoke.log; tail -n 4 ${P}_bench_reference.err ${P}_bench_ours.err
python - <<'PY'
import json
for f in ("gpurun_out/r3f_bench_reference.json","gpurun_out/r3f_bench_ours.json"):
    try:
        d=json.loads(open(f).read())
        print(f, round(d.get("value",0)), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_checked"), (d.get("full_model") or {}).get("images_per_sec"), (d.get("cpu_baseline") or {}).get("value"), d.get("hbm_frac_whole_step"))
        for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("images_per_sec"), v.get("hbm_frac_whole_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
