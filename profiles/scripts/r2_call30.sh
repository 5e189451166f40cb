mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=30 2>&1 | tail -5 | cut -c1-300
python profiles/microbench_glue.py 2>&1 | grep -i instnorm
HDF_IN_FLOAT_STAGED=1 python profiles/microbench_glue.py 2>&1 | grep -i instnorm
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c30_bench.json')); print('raw staged', d['value'], d['ms_per_step'])"
HDF_IN_FLOAT_STAGED=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c30_bench_old.json 2> gpurun_out/c30_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c30_bench_old.json')); print('float staged', d['value'], d['ms_per_step'])"
done
