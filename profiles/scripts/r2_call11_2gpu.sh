mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -150 > gpurun_out/c11_pytest_multi.txt; grep -v "^E   *$" gpurun_out/c11_pytest_multi.txt | tail -70
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c11_bench_2gpu.json 2> gpurun_out/c11_bench_2gpu.err; head -c 3000 gpurun_out/c11_bench_2gpu.json; tail -5 gpurun_out/c11_bench_2gpu.err
