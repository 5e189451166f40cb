mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --maxfail=30 -k "stem" 2>&1 | tail -5 | cut -c1-300
python profiles/microbench_stem.py 2>&1 | tee gpurun_out/r2_microbench_stem.txt
HDF_STEM_TILES_PER_CTA=0 python profiles/microbench_stem.py 2>&1 | head -1
for t in 32 16 64 0; do
HDF_STEM_TILES_PER_CTA=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c21_bench_$t.json 2> gpurun_out/c21_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c21_bench_$t.json')); print('tiles/cta $t', d['value'], d['ms_per_step'])"
done
