#!/bin/bash
# Round-2 second GPU visit: parity of the resident single-launch forward, tanh detail, micro-benchmarks, resident A/B,
# e2e host-copy thread sweep, single-pass ncu of an in-step step (dram bytes per launch, caches left as the step leaves them).
set -u
mkdir -p gpurun_out
P=gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_selftest.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -40 > ${P}_pytest_new.log
python - > ${P}_tanh_detail.log 2>&1 <<'PY'
import ctypes, struct
import b200quant
from b200quant import _lib
ctx = _lib.context(0)
for which in (1, 2, 3, 4):
    n = ctypes.c_int64(-1)
    ctx.call("b2q_selftest", which, ctypes.byref(n))
    v = n.value
    extra = ""
    if which in (3, 4) and v:
        x = struct.unpack("f", struct.pack("I", v))[0]
        extra = " x=%r bits=0x%08x" % (x, v)
    print("selftest", which, v, extra)
PY
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -30 > ${P}_pytest_fullsize.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_edge_cases.py --deselect tests/test_gpu_selftest.py 2>&1 | tail -40 > ${P}_pytest_rest.log
B2Q_OPT_RESIDENT=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -8 > ${P}_pytest_resident0.log
timeout 600 python tools/microbench2.py > ${P}_microbench2.log 2>&1
timeout 1200 python bench.py --steps 10 --warmup 5 > ${P}_bench_ours.json 2> ${P}_bench_ours.err
SHORT="--steps 10 --warmup 5 --no-e2e --no-cpu --no-workloads --no-micro --no-full-model"
B2Q_OPT_RESIDENT=0 timeout 600 python bench.py $SHORT > ${P}_bench_resident0.json 2> ${P}_bench_resident0.err
B2Q_OPT_RESIDENT_MAX_MB=110 timeout 600 python bench.py $SHORT > ${P}_bench_resident110.json 2> ${P}_bench_resident110.err
B2Q_OPT_RESIDENT_MAX_MB=30 timeout 600 python bench.py $SHORT > ${P}_bench_resident30.json 2> ${P}_bench_resident30.err
E2E="--steps 3 --warmup 3 --no-cpu --no-workloads --no-micro --no-full-model --no-graph --e2e-steps 3"
for T in 2 4 8 12; do
  B2Q_HOST_COPY_THREADS=$T timeout 600 python bench.py $E2E > ${P}_e2e_threads$T.json 2> ${P}_e2e_threads$T.err
done
B2Q_OPT_HOST_STE_COPY=0 timeout 600 python bench.py $E2E > ${P}_e2e_gpucopy.json 2> ${P}_e2e_gpucopy.err
PROF="python bench.py --steps 1 --warmup 3 --profile"
B2Q_OPT_RESIDENT=0 $PROF > ${P}_plain0.log 2>&1 &&
B2Q_OPT_RESIDENT=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 972 -c 330 --csv --log-file ${P}_instep_resident0.csv $PROF > ${P}_ncu0.log 2>&1
$PROF > ${P}_plain1.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 850 -c 330 --csv --log-file ${P}_instep_resident1.csv $PROF > ${P}_ncu1.log 2>&1
for f in ${P}_pytest_new.log ${P}_pytest_fullsize.log ${P}_pytest_rest.log ${P}_pytest_resident0.log; do tail -n 3 $f; done
cat ${P}_tanh_detail.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_bench_*.json"))+sorted(glob.glob("gpurun_out/r2b_e2e_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f, d.get("ms_per_step_by_mode"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
