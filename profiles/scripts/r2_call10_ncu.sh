mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "metric_tail or token_tensor" 2>&1 | tail -5 > gpurun_out/c10_pytest.txt; cat gpurun_out/c10_pytest.txt
# 1. launch list of one eager steady-state step (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_v1.csv python profiles/one_step.py > gpurun_out/c10_one_step.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v1.csv 60 > gpurun_out/r2_launches_v1.txt 2>&1; head -30 gpurun_out/r2_launches_v1.txt
# 2. ncu --set full of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_ws_kernel -c 2 -o gpurun_out/r2_ncu_ws_b11r -f python profiles/microbench_conv.py --reps 1 --only block_1_1_right > gpurun_out/c10_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_ws_kernel -c 1 -o gpurun_out/r2_ncu_ws_b12l -f python profiles/microbench_conv.py --reps 1 --only block_1_2_left > gpurun_out/c10_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_ws_kernel -c 1 -o gpurun_out/r2_ncu_wgrad_ws_32x32 -f python profiles/microbench_conv.py --reps 1 --only block_1_2_left > gpurun_out/c10_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_fwd_kernel -c 1 -o gpurun_out/r2_ncu_fwd_b21r -f python profiles/microbench_conv.py --reps 1 --only block_2_1_right > gpurun_out/c10_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
