#!/bin/bash
# Round-2 first GPU visit: all GPU tests (new edge / full-size / selftest files first), micro-benchmarks, both bench arms.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_selftest.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2a_pytest_new.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=10 2>&1 | tail -60 > gpurun_out/r2a_pytest_fullsize.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_edge_cases.py --deselect tests/test_gpu_selftest.py 2>&1 | tail -40 > gpurun_out/r2a_pytest_rest.log
timeout 600 python tools/microbench2.py > gpurun_out/r2a_microbench2.log 2>&1
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
timeout 1200 python bench.py --steps 10 --warmup 5 > gpurun_out/r2a_bench_ours.json 2> gpurun_out/r2a_bench_ours.err
tail -5 gpurun_out/r2a_pytest_new.log gpurun_out/r2a_pytest_fullsize.log gpurun_out/r2a_pytest_rest.log
tail -c 1500 gpurun_out/r2a_bench_ours.err
head -c 1500 gpurun_out/r2a_bench_ours.json
