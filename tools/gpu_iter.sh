#!/bin/bash
# quick iteration: op tests, unet tests, per-kernel timing, bench
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -12 ) | tee gpurun_out/ops.log
( timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -12 ) | tee gpurun_out/unet_tests.log
( timeout 300 python tools/profile_unet.py --kernels --shapes 2>&1 | tail -75 ) | tee gpurun_out/shapes.log
( timeout 400 python bench.py --steps 40 --warmup 3 2>&1 | tail -3 ) | tee gpurun_out/bench.log
