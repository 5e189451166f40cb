mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -15 > gpurun_out/c22_pytest.txt
tail -6 gpurun_out/c22_pytest.txt | cut -c1-300
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c22_bench.json')); print('fused finalize', d['value'], d['ms_per_step'], d['gpu_launches'])"
HDF_NO_FUSED_FINALIZE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c22_bench_nofin.json 2> gpurun_out/c22_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c22_bench_nofin.json')); print('separate finalize', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
