mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model2d.py -m gpu -q --maxfail=30 2>&1 | tail -40 | cut -c1-300 > gpurun_out/c38_pytest2d.txt; tail -40 gpurun_out/c38_pytest2d.txt
