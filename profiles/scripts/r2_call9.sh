mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -q --maxfail=30 2>&1 | tail -8 > gpurun_out/c9_pytest.txt
tail -4 gpurun_out/c9_pytest.txt
python profiles/microbench_conv.py --reps 5 > gpurun_out/r2_microbench_conv_v4_prewait.txt 2>&1
cat gpurun_out/r2_microbench_conv_v4_prewait.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c9_bench.json')); print(d['value'], d['ms_per_step'])"
