mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prep.py tests/test_gpu_bench_config.py -m gpu -q --maxfail=30 2>&1 | tail -40 > gpurun_out/c17_pytest.txt
tail -30 gpurun_out/c17_pytest.txt
python profiles/bench_input_pipeline.py > gpurun_out/r2_input_pipeline.json 2> gpurun_out/c17_pipe.err; cat gpurun_out/r2_input_pipeline.json; tail -3 gpurun_out/c17_pipe.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c17_bench.json')); print(d['value'], d['ms_per_step'])"
