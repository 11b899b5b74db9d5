#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8 ) | tee gpurun_out/tests.log
( timeout 400 python bench.py --steps 40 --warmup 3 --no-batched 2>&1 | tail -1 ) | tee gpurun_out/bench.log | cut -c1-200
