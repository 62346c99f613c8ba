python -m pytest tests/test_gpu_bnfold.py -m gpu -q -x -k special 2>&1 | grep -E "Error|assert|Mismatch|Max|x:|y:|^E" | head -30
