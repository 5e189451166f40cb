mkdir -p gpurun_out
# 1. launch list of one eager steady-state step of the final code (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_v2.csv python profiles/one_step.py > gpurun_out/c23_one_step.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v2.csv 60 > gpurun_out/r2_launches_v2.txt 2>&1; head -24 gpurun_out/r2_launches_v2.txt
# 2. ncu --set full of the kernels added since call 10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_convt_kernel -c 1 -o gpurun_out/r2_ncu_convt_upconv1 -f python profiles/microbench_conv.py --reps 1 --only "upconv_1(T)" > gpurun_out/c23_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tc_fwd_kernel -c 1 -o gpurun_out/r2_ncu_stem_fwd -f python profiles/microbench_stem.py > gpurun_out/c23_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tc_wgrad_kernel -c 1 -o gpurun_out/r2_ncu_stem_wgrad -f python profiles/microbench_stem.py > gpurun_out/c23_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upsample2_bwd_blk_kernel -c 1 -o gpurun_out/r2_ncu_ups_bwd -f python profiles/microbench_glue.py > gpurun_out/c23_ncu4.log 2>&1
for f in convt_upconv1 stem_fwd stem_wgrad ups_bwd; do python profiles/ncu_summary.py gpurun_out/r2_ncu_$f.ncu-rep > gpurun_out/r2_ncu_$f.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
