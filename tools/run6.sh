#!/bin/bash
mkdir -p gpurun_out
python tools/halo_timeline.py > gpurun_out/timeline.log 2>&1
for f in 1 2 3 0; do B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick6.log 2>&1; done
B200SD_FUSED=3 B200SD_HALO_MIN_HW=4096 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick6.log 2>&1
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_factory_gpu.py tests/test_rng.py -q -m gpu > gpurun_out/fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/fused_tests.log
B200SD_FUSED=2 timeout 900 python -m pytest tests/test_unet_gpu.py -q -x -k "tiny_vs_oracle or sd21_base or graph_equals" > gpurun_out/unet_apply_tests.log 2>&1; echo "rc=$?" >> gpurun_out/unet_apply_tests.log
cat gpurun_out/timeline.log; grep quick gpurun_out/quick6.log; tail -8 gpurun_out/fused_tests.log; tail -4 gpurun_out/unet_apply_tests.log
