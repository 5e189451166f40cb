mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -4 | cut -c1-300
for i in 1 2 3; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c37_bench.json')); print('apply+head', d['value'], d['ms_per_step'])" || tail -3 gpurun_out/c37_bench.err
HDF_NO_APPLY_HEAD=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c37_bench_old.json 2> gpurun_out/c37_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c37_bench_old.json')); print('separate', d['value'], d['ms_per_step'])"
done
