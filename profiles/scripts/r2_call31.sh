mkdir -p gpurun_out
for p in 0 -1 -2 -3; do
HDF_BWD_SIDE_PRIO=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window --no-input-pipeline > gpurun_out/c31_bench.json 2> gpurun_out/c31_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c31_bench.json')); print('bwd side prio $p', d['value'], d['ms_per_step'])"
done
HDF_BWD_SIDE_PRIO=-2 HDF_TL_FIRST=4 python profiles/timeline_overlap.py > gpurun_out/r2_timeline_v8_prio.txt 2>&1; grep -E "attn_bwd first|last tc_conv_wgrad|adam start|idle total" gpurun_out/r2_timeline_v8_prio.txt | cut -c1-200
