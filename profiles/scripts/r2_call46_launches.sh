mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_v3.csv python profiles/one_step.py > gpurun_out/c46_one_step.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v3.csv 60 > gpurun_out/r2_launches_v3.txt 2>&1; head -16 gpurun_out/r2_launches_v3.txt
