mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
( cd profiles/probes && timeout 300 ./umma_probe2.bin > ../../gpurun_out/r2_umma_probe2.txt 2>&1; echo "probe exit $?" >> ../../gpurun_out/r2_umma_probe2.txt )
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/c1_pytest.txt
bash profiles/scripts/run_sanitizer.sh initcheck racecheck synccheck > gpurun_out/c1_sanitizer.txt 2>&1
tail -5 gpurun_out/c1_pytest.txt; tail -30 gpurun_out/r2_umma_probe2.txt
