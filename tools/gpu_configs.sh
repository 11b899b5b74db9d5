#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x -k "causal or text_encoder or attention" 2>&1 | tail -15 ) | tee gpurun_out/tests_text.log
