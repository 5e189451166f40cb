mkdir -p gpurun_out
HDF_GLUE_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"head_wgrad_kernel|head_dgrad_fast_kernel|head_fwd_kernel|in_bwd_apply_fast_kernel|rowreduce_kernel|in_apply_fast_kernel|maxpool_bwd" -c 30 -o gpurun_out/r2_ncu_glue -f python profiles/microbench_glue.py > gpurun_out/c29_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_ncu_glue.ncu-rep > gpurun_out/r2_ncu_glue.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes_read|dram__bytes_write|gpu__dram_throughput|sm__warps_active|registers|grid_size|sm__throughput" gpurun_out/r2_ncu_glue.txt | cut -c60-175
