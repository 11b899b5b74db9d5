#!/bin/bash
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "pipeline_tiny_end or xl_text" 2>&1 | tail -4 ) | tee gpurun_out/tests_new.log
