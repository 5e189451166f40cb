mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=30 2>&1 | tail -15 > gpurun_out/c7_pytest.txt
tail -6 gpurun_out/c7_pytest.txt
HDF_TL_FIRST=60 HDF_TL_WINDOWS="14.0:14.1" python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v3.txt 2>&1; sed -n 8,24p gpurun_out/r2_timeline_v3.txt; grep -E "tok_a_fwd_kernel|tok_c_fwd_kernel" gpurun_out/r2_timeline_v3.txt | head -12
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c7_bench.json')); print(d['value'], d['ms_per_step'], d['sliding_window'])"
