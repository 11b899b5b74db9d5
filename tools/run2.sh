#!/bin/bash
# gpurun round-2 call #2: fused-normalisation op tests (one process per test: a hung kernel cannot block the rest),
# then the UNet parity tests on the fused graph and a fused / unfused quick bench
mkdir -p gpurun_out
rm -f gpurun_out/fused_tests.log
for t in test_halo_conv_plain test_halo_conv_two_sources_temb_residual test_halo_conv_groupnorm_silu test_halo_upsample_conv \
         test_conv_column_statistics test_linear_staged_residual_row_and_column_statistics test_layernorm_folded_into_linear \
         test_halo_1x1_groupnorm_rowstats; do
  echo "=== $t" >> gpurun_out/fused_tests.log
  timeout 240 python -m pytest tests/test_fused_gpu.py -q -x -k "$t" 2>&1 | tail -15 >> gpurun_out/fused_tests.log
  echo "rc=$?" >> gpurun_out/fused_tests.log
done
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x > gpurun_out/ops_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ops_tests.log
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -x > gpurun_out/unet_tests.log 2>&1; echo "rc=$?" >> gpurun_out/unet_tests.log
for f in 1 0; do
  B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick2.log 2>&1
done
grep -E "^===|passed|failed|rc=|Error|error" gpurun_out/fused_tests.log | head -60
tail -5 gpurun_out/ops_tests.log; tail -15 gpurun_out/unet_tests.log; grep quick gpurun_out/quick2.log
