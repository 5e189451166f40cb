mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=30 2>&1 | tail -30 > gpurun_out/c15_pytest.txt
tail -12 gpurun_out/c15_pytest.txt
python profiles/microbench_conv.py --reps 5 > gpurun_out/r2_microbench_conv_v5_convt.txt 2>&1
tail -6 gpurun_out/r2_microbench_conv_v5_convt.txt
python profiles/microbench_glue.py > gpurun_out/r2_microbench_glue_v1.txt 2>&1; grep -i "upsample\|maxpool" gpurun_out/r2_microbench_glue_v1.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c15_bench.json')); print(d['value'], d['ms_per_step'])"
HDF_TC_NO_CONVT=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c15_bench_noconvt.json 2> gpurun_out/c15_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c15_bench_noconvt.json')); print('no convt', d['value'], d['ms_per_step'])"
HDF_UPS_BWD_GATHER=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window > gpurun_out/c15_bench_gather.json 2> gpurun_out/c15_bench3.err; python -c "
import json; d=json.load(open('gpurun_out/c15_bench_gather.json')); print('ups gather', d['value'], d['ms_per_step'])"
