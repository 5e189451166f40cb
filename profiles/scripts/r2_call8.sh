mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -8 > gpurun_out/c8_pytest.txt
tail -4 gpurun_out/c8_pytest.txt
HDF_TL_FIRST=10 HDF_TL_WINDOWS="14.0:14.05" python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v4.txt 2>&1; sed -n 8,64p gpurun_out/r2_timeline_v4.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c8_bench.json')); print(d['value'], d['ms_per_step'])"
