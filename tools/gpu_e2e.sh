#!/bin/bash
# UNet / VAE / pipeline parity on the GPU, smoke(), then the bench line.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -s 2>&1 | tail -40 ) | tee gpurun_out/unet_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/smoke.log
( timeout 600 python bench.py --steps 40 --warmup 3 2>&1 | tail -5 ) | tee gpurun_out/bench.log
