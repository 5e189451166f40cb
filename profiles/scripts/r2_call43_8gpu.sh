mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/c43_bench_8gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_8gpu_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sliding_window']['ms_per_volume'], d['sliding_window']['dice_vs_oracle_fp32_mask'])
PY
tail -n 2 gpurun_out/c43_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 4 --steps 10 --warmup 3 --no-sw-dice > gpurun_out/r2_bench_4gpu_final.json 2> gpurun_out/c43_bench_4gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_4gpu_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sliding_window']['ms_per_volume'])
PY
