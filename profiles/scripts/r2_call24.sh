mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -8 > gpurun_out/c24_pytest.txt
tail -4 gpurun_out/c24_pytest.txt | cut -c1-300
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c24_bench.json')); print('tiled reduce', d['value'], d['ms_per_step'])"
HDF_TC_WGRAD_REDUCE_SIMPLE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c24_bench_simple.json 2> gpurun_out/c24_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c24_bench_simple.json')); print('simple reduce', d['value'], d['ms_per_step'])"
done
