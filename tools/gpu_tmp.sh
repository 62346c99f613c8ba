timeout 600 python -m pytest tests/test_gpu_bnfold.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/bn_sweep.py 2>&1 | tail -8
