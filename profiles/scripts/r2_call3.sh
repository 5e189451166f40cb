mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --maxfail=30 2>&1 | tail -15 > gpurun_out/c3_pytest.txt
tail -5 gpurun_out/c3_pytest.txt
python profiles/microbench_conv.py --reps 5 --only block_1 > gpurun_out/r2_microbench_conv_v2_ws.txt 2>&1
python profiles/microbench_conv.py --reps 5 --only up3 >> gpurun_out/r2_microbench_conv_v2_ws.txt 2>&1
HDF_TC_DEBUG=1 python profiles/microbench_conv.py --reps 1 --only block_1_2_left 2>&1 | grep -v "^layer\|^TOTAL" | tail -3 > gpurun_out/c3_dbg.txt
HDF_TC_DEBUG=1 python profiles/microbench_conv.py --reps 1 --only block_1_1_right 2>&1 | grep -v "^layer\|^TOTAL" | tail -4 >> gpurun_out/c3_dbg.txt
cat gpurun_out/r2_microbench_conv_v2_ws.txt gpurun_out/c3_dbg.txt
python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v1.txt 2>&1; tail -40 gpurun_out/r2_timeline_v1.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; cat gpurun_out/c3_bench.json | head -c 6000; tail -5 gpurun_out/c3_bench.err
