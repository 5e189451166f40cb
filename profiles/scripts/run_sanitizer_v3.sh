#!/bin/bash
# compute-sanitizer, third pass: code added after profiles/sanitizer/r2b_summary.txt -- the 2-D model on flat volumes (depth-factor
# pool / up-sample / patch / loss entries, lifted transposed conv, tap filtering), the InstanceNorm-apply + head fusion, the early
# weight-gradient stream and the token weight-gradient streams (exercised by the bf16 model step).
OUT=gpurun_out/sanitizer
mkdir -p $OUT
TOOLS=${@:-"memcheck racecheck synccheck initcheck"}
SEL='test_fp32_forward_backward_matches_reference_golden_2d or (test_fp32_and_bf16_vs_oracle_2d and 3-2-32) or (test_instnorm_apply_fused_with_head and (32-2-4097 or 256-3-77)) or (test_bf16_tensor_core_path_other_configs and 3-2-size0)'
for tool in $TOOLS; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 --log-file $OUT/r2c_${tool}.log \
    python -m pytest tests/test_gpu_model.py tests/test_gpu_model2d.py tests/test_gpu_ops.py -q -k "$SEL" > $OUT/r2c_${tool}.pytest.txt 2>&1
  echo "exit $?" >> $OUT/r2c_${tool}.pytest.txt
  tail -3 $OUT/r2c_${tool}.pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r2c_${tool}.log | tail -2
done
