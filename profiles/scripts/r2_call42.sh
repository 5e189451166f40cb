mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model2d.py tests/test_gpu_tc.py -m gpu -q --maxfail=30 2>&1 | grep -E "Error|error|passed|failed" | tail -6 | cut -c1-300
python profiles/bench_2d.py > gpurun_out/r2_bench_2d.json 2> gpurun_out/c42.err; cat gpurun_out/r2_bench_2d.json | cut -c1-420
