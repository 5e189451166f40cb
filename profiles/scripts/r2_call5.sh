mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -40 > gpurun_out/c5_pytest.txt
tail -25 gpurun_out/c5_pytest.txt
python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v2.txt 2>&1; head -24 gpurun_out/r2_timeline_v2.txt | tail -20
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sliding-window --no-eager-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c5_bench.json')); print(d['value'], d['ms_per_step'])"
HDF_NO_TOK_TC=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sliding-window --no-eager-baseline > gpurun_out/c5_bench_notok.json 2> gpurun_out/c5_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c5_bench_notok.json')); print('no tok tc', d['value'], d['ms_per_step'])"
