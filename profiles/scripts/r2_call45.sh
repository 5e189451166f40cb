timeout 600 python -m pytest tests/test_gpu_prep.py -m gpu -q 2>&1 | tail -2
