mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -150 > gpurun_out/c27_pytest_multi.txt; grep -v "^E   *$" gpurun_out/c27_pytest_multi.txt | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/c27_bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sliding_window'])
PY
tail -n 3 gpurun_out/c27_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/c27_ref2.err | cut -c1-300
