#!/bin/bash
# compute-sanitizer over the 32^3 model step (fp32 exact path + bf16 tensor-core path) and the three tcgen05 kernels
# (VERDICT r1 item 1d).  Run on the GPU box: bash profiles/scripts/run_sanitizer.sh [tools...]
OUT=gpurun_out/sanitizer
mkdir -p $OUT
TOOLS=${@:-"initcheck racecheck synccheck memcheck"}
SEL='model_nf16_32cube or (test_tc_conv_fwd_matches_torch and (32-32-size1 or 64-32-size4)) or (test_tc_conv_wgrad_matches_torch and 32-32) or (test_tc_conv_transpose and 64-32-size1) or test_bf16_tensor_core_path_other_configs'
for tool in $TOOLS; do
  echo "=== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 --log-file $OUT/r2_${tool}.log \
    python -m pytest tests/test_gpu_model.py tests/test_gpu_tc.py -x -q -k "$SEL" > $OUT/r2_${tool}.pytest.txt 2>&1
  echo "exit $?" >> $OUT/r2_${tool}.pytest.txt
  tail -3 $OUT/r2_${tool}.pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r2_${tool}.log | tail -2
done
