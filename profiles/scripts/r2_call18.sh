mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prep.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 -k "prep or mr_configs" 2>&1 | tail -40 > gpurun_out/c18_pytest.txt
tail -12 gpurun_out/c18_pytest.txt | cut -c1-600
python profiles/bench_input_pipeline.py > gpurun_out/r2_input_pipeline.json 2> gpurun_out/c18_pipe.err; cat gpurun_out/r2_input_pipeline.json; tail -3 gpurun_out/c18_pipe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_parity_table.json'))
for k,v in d.items():
    if k.startswith('config3') or k.startswith('config4'): print(k, json.dumps(v))
PY
