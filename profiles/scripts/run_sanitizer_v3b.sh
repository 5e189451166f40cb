#!/bin/bash
# initcheck re-run of the third pass WITHOUT the test that also runs the reference graph through torch.autocast on the GPU: the
# 2.4 M "uninitialized read" reports of r2c_initcheck.log are all in torch's own cast kernel reading the bf16 output of a cuBLASLt
# GEMM (layer_norm under autocast in the oracle yardstick) -- stores issued through TMA are not tracked by initcheck.
OUT=gpurun_out/sanitizer
mkdir -p $OUT
SEL='test_fp32_forward_backward_matches_reference_golden_2d or test_2d_graphed_train_step_matches_eager or (test_instnorm_apply_fused_with_head and (32-2-4097 or 256-3-77)) or (test_bf16_tensor_core_path_other_configs and 3-2-size0)'
timeout 900 compute-sanitizer --tool initcheck --print-limit 100 --error-exitcode 0 --log-file $OUT/r2c_initcheck_ours.log \
  python -m pytest tests/test_gpu_model.py tests/test_gpu_model2d.py tests/test_gpu_ops.py -q -k "$SEL" > $OUT/r2c_initcheck_ours.pytest.txt 2>&1
echo "exit $?" >> $OUT/r2c_initcheck_ours.pytest.txt
tail -3 $OUT/r2c_initcheck_ours.pytest.txt
grep -E "ERROR SUMMARY" $OUT/r2c_initcheck_ours.log | tail -2
grep -E "^=========     at " $OUT/r2c_initcheck_ours.log | cut -c1-160 | sort | uniq -c | head
