#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x -k "pad_after or vae_encoder or sdxl_outputs" 2>&1 | tail -15 ) | tee gpurun_out/tests_new.log
