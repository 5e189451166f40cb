mkdir -p gpurun_out
HDF_TL_FIRST=150 HDF_TL_WINDOWS="14.0:14.4" python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v2.txt 2>&1; head -75 gpurun_out/r2_timeline_v2.txt | tail -68
