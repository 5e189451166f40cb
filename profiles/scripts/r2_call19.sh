mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -40 > gpurun_out/c19_pytest.txt
tail -12 gpurun_out/c19_pytest.txt | cut -c1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c19_bench.json')); print(d['value'], d['ms_per_step'])"
HDF_TL_FIRST=14 python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v6.txt 2>&1; grep -E "patch|PatchA|gemm_tile|tok_a_fwd first" gpurun_out/r2_timeline_v6.txt | head; sed -n 70,90p gpurun_out/r2_timeline_v6.txt
