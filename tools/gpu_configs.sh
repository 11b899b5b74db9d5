#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python tools/bench_gemm_shapes.py 2>&1 | tail -40 ) | tee gpurun_out/gemm_shapes.log
( timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "controlnet" 2>&1 | tail -8 ) | tee gpurun_out/tests.log
