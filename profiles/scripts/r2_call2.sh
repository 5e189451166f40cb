mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -60 > gpurun_out/c2_pytest.txt
tail -15 gpurun_out/c2_pytest.txt
python profiles/microbench_conv.py --reps 5 > gpurun_out/r2_microbench_conv_v1_ws.txt 2>&1
HDF_TC_NO_WS=1 python profiles/microbench_conv.py --reps 5 --only block_1 > gpurun_out/r2_microbench_conv_v1_nows.txt 2>&1
HDF_TC_NO_WS=1 python profiles/microbench_conv.py --reps 5 --only up3 >> gpurun_out/r2_microbench_conv_v1_nows.txt 2>&1
HDF_TC_DEBUG=1 python profiles/microbench_conv.py --reps 1 --only block_1_2_left > gpurun_out/c2_dbg.txt 2>&1
HDF_TC_DEBUG=1 python profiles/microbench_conv.py --reps 1 --only block_1_1_right >> gpurun_out/c2_dbg.txt 2>&1
HDF_TC_DEBUG=1 python profiles/microbench_conv.py --reps 1 --only block_2_2_left >> gpurun_out/c2_dbg.txt 2>&1
cat gpurun_out/r2_microbench_conv_v1_ws.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; cat gpurun_out/c2_bench.json | head -c 1500
