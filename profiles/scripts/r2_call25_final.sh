mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# the driver's two arms, default flags
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r2_bench_final_reference_arm.json 2> gpurun_out/c25_ref.err; cat gpurun_out/r2_bench_final_reference_arm.json | cut -c1-600; grep real gpurun_out/c25_ref.err
( time python bench.py ) > gpurun_out/r2_bench_final.json 2> gpurun_out/c25_bench.err; grep real gpurun_out/c25_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'sw', d['sliding_window'], )
print('pipe', d['input_pipeline']); print('eager', d['gpu_eager_baseline']); print('cpu', d.get('cpu_baseline'))
print('roofline classes', d['roofline'].get('classes')); print({k:v for k,v in d['roofline'].items() if k!='classes'})
PY
# other BASELINE configs
python bench.py --depth 24 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/r2_bench_config2_td24.json 2>> gpurun_out/c25_cfg.err
python bench.py --modalities 3 --size 32 384 384 --batch 1 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/r2_bench_config3_3mod_32x384x384.json 2>> gpurun_out/c25_cfg.err
python bench.py --modalities 4 --classes 4 --size 128 128 128 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/r2_bench_config4_4mod_128.json 2>> gpurun_out/c25_cfg.err
python - <<'PY'
import json
for f in ['r2_bench_config2_td24','r2_bench_config3_3mod_32x384x384','r2_bench_config4_4mod_128']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['value'], d['ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/c25_cfg.err
