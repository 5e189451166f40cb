#!/bin/bash
# compute-sanitizer over the kernels added after the first round-2 pass (profiles/sanitizer/r2_summary.txt): weight-stationary
# conv + fused statistics, plane-ring weight gradient, shift-major transposed conv, fused first conv (forward + weight
# gradient), tcgen05 patch embedding, token tensor-core forward, input pipeline, plus the 32^3 bf16 model step that strings
# them together.  Run on the GPU box: bash profiles/scripts/run_sanitizer_v2.sh [tools...]
OUT=gpurun_out/sanitizer
mkdir -p $OUT
TOOLS=${@:-"memcheck racecheck synccheck initcheck"}
SEL='(test_tc_ws_conv_matches_torch_and_old_kernel and (32-size0 or 64-size1)) or (test_shift_major_transposed_conv and (size1 or size2)) or (test_stem_fused_gather and (2-32-size0 or 4-32-size3 or 1-16-size1)) or (test_patch_embed_tcgen05 and 1-size1-64) or (test_token_tensor_core_layer and 2-37-160) or test_pipeline_matches_reference_golden or test_bf16_tensor_core_path_other_configs or (test_tc_conv_wgrad_matches_torch and 32-32-size1-2)'
for tool in $TOOLS; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 --log-file $OUT/r2b_${tool}.log \
    python -m pytest tests/test_gpu_model.py tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_prep.py -q -k "$SEL" > $OUT/r2b_${tool}.pytest.txt 2>&1
  echo "exit $?" >> $OUT/r2b_${tool}.pytest.txt
  tail -3 $OUT/r2b_${tool}.pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r2b_${tool}.log | tail -2
done
