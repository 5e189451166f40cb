mkdir -p gpurun_out
python profiles/bench_2d.py > gpurun_out/r2_bench_2d.json 2> gpurun_out/c40.err; cat gpurun_out/r2_bench_2d.json; tail -3 gpurun_out/c40.err
