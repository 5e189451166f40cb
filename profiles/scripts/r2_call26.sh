mkdir -p gpurun_out
python bench.py --modalities 3 --size 32 384 384 --batch 1 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/r2_bench_config3_3mod_32x384x384.json 2> gpurun_out/c26_cfg3.err
python bench.py --modalities 4 --classes 4 --size 128 128 128 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/r2_bench_config4_4mod_128.json 2> gpurun_out/c26_cfg4.err
python - <<'PY'
import json
for f in ['r2_bench_config3_3mod_32x384x384','r2_bench_config4_4mod_128']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/c26_cfg3.err gpurun_out/c26_cfg4.err
