mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model2d.py tests/test_gpu_tc.py -m gpu -q --maxfail=30 -k "2d or ws" 2>&1 | tail -4 | cut -c1-300
python profiles/bench_2d.py > gpurun_out/r2_bench_2d.json 2> gpurun_out/c41.err; cat gpurun_out/r2_bench_2d.json | cut -c1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c41_bench.json')); print('3D bench', d['value'], d['ms_per_step'])"
