mkdir -p gpurun_out
for k in 0 1 2 4 8; do
HDF_WGRAD_EARLY=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c32_bench.json')); print('wgrad early $k', d['value'], d['ms_per_step'])" || tail -3 gpurun_out/c32_bench.err
done
HDF_WGRAD_EARLY=2 timeout 600 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -k "graphed_train_step or config1" 2>&1 | tail -3
