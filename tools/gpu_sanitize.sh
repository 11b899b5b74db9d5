#!/bin/bash
# compute-sanitizer memcheck over the operator-level GPU tests (GEMM / convolution variants incl. folded shortcut and halo
# kernels, attention incl. stream-K, GroupNorm) and the tiny-UNet tests through the default fused path
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest -q -m gpu \
  tests/test_ops_gpu.py tests/test_fused_gpu.py "tests/test_unet_gpu.py::test_unet_tiny_vs_oracle_and_golden[ORIGINAL]" \
  tests/test_unet_gpu.py::test_unet_tiny_controlnet_residuals tests/test_unet_gpu.py::test_pipeline_tiny_end_to_end_vs_oracle \
  tests/test_capi_gpu.py::test_capi_unet_tiny_matches_python_engine_and_oracle \
  > gpurun_out/sanitize.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize.log
grep "ERROR SUMMARY\|passed\|failed\|rc=" gpurun_out/sanitize.log | tail -6; grep -m8 "Invalid\|Error" gpurun_out/sanitize.log; true
