mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -6 | cut -c1-300
