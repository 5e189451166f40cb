mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -12 > gpurun_out/c12_pytest.txt; tail -6 gpurun_out/c12_pytest.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c12_bench.json')); print(d['value'], d['ms_per_step'], d['sliding_window'])"
HDF_NO_FUSED_STATS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c12_bench_nostats.json 2> gpurun_out/c12_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c12_bench_nostats.json')); print('no fused stats', d['value'], d['ms_per_step'])"
python profiles/microbench_conv.py --reps 5 --only block_1 > gpurun_out/c12_mb.txt 2>&1; cat gpurun_out/c12_mb.txt
