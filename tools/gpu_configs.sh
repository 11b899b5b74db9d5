#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "image_to_image or pipeline_tiny_end" 2>&1 | tail -15 ) | tee gpurun_out/tests_new.log
