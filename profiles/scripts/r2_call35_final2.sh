mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -6 > gpurun_out/c35_pytest.txt; tail -4 gpurun_out/c35_pytest.txt | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/r2_bench_final.json 2> gpurun_out/c35_bench.err; grep real gpurun_out/c35_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'sw', d['sliding_window']['ms_per_volume'], d['sliding_window']['dice_vs_oracle_fp32_mask'])
print('pipe', d['input_pipeline']['ms_per_batch']); print('eager', d['gpu_eager_baseline']['value']); print('cpu', d.get('cpu_baseline'))
print('roofline', d['roofline']['frac'], {k: (round(v['ms'],3), round(v['frac'],3)) for k,v in d['roofline']['classes'].items()}, d['roofline']['step_frac_of_sustained_peak'])
PY
HDF_TL_FIRST=4 python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v10_final.txt 2>&1
