mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c16_bench.json')); print('default', d['value'], d['ms_per_step'], d['clocks'])"
HDF_TC_PACK_SIMPLE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c16_bench_simplepack.json 2> gpurun_out/c16_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c16_bench_simplepack.json')); print('simple pack', d['value'], d['ms_per_step'])"
done
HDF_TL_FIRST=10 HDF_TL_WINDOWS="14.0:14.05" python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v5.txt 2>&1; sed -n 8,70p gpurun_out/r2_timeline_v5.txt
