mkdir -p gpurun_out
HDF_TL_FIRST=4 HDF_TL_WINDOWS="19.0:21.5" python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v9.txt 2>&1; grep -E "span ms|union busy|attn_bwd first|last tc_conv_wgrad|adam start|idle total" gpurun_out/r2_timeline_v9.txt | cut -c1-250; grep -n "kernels starting in" -A60 gpurun_out/r2_timeline_v9.txt | cut -c1-130 | tail -70
