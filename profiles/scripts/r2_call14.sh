mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sliding-window"
run() { name=$1; shift; env "$@" $B > gpurun_out/c14_$name.json 2> gpurun_out/c14_$name.err; python -c "
import json,sys; d=json.load(open('gpurun_out/c14_$name.json')); print('$name', round(d['value'],2), round(d['ms_per_step'],3))"; }
run base X=1
run nopatchfirst HDF_NO_PATCH_FIRST=1
run nmax128 HDF_TC_WS_NMAX=128
run stages5 HDF_TC_WS_STAGES=5
run nodefer HDF_NO_DEFER_WGRAD=1
run fwdprio0 HDF_FWD_SIDE_PRIO=0
run base2 X=1
