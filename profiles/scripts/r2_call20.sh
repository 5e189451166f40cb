mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --maxfail=30 -k "stem" 2>&1 | tail -30 > gpurun_out/c20_pytest_stem.txt
tail -15 gpurun_out/c20_pytest_stem.txt | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -30 > gpurun_out/c20_pytest.txt
tail -8 gpurun_out/c20_pytest.txt | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c20_bench.json')); print(d['value'], d['ms_per_step'])"
HDF_NO_STEM_FUSED=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c20_bench_old.json 2> gpurun_out/c20_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c20_bench_old.json')); print('im2col stem', d['value'], d['ms_per_step'])"
HDF_TL_FIRST=14 python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v7.txt 2>&1; grep -E "stem|idle total" gpurun_out/r2_timeline_v7.txt | head
