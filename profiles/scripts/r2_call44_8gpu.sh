mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for v in hi default; do
if [ $v = default ]; then export HDF_NCCL_DEFAULT_PRIORITY=1; else unset HDF_NCCL_DEFAULT_PRIORITY; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 10 --warmup 3 --no-sliding-window > gpurun_out/c44_bench_8gpu_$v.json 2> gpurun_out/c44_$v.err; python - <<PY
import json
d=json.load(open('gpurun_out/c44_bench_8gpu_$v.json')); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
