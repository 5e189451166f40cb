mkdir -p gpurun_out
for k in 8 10 12 14 16 20; do
HDF_WGRAD_EARLY=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c33_bench.json')); print('wgrad early $k', d['value'], d['ms_per_step'])" || tail -3 gpurun_out/c33_bench.err
done
