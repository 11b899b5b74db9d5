#!/bin/bash
mkdir -p gpurun_out
( timeout 45 python tools/check_controlnet_golden.py 2>&1 | tail -15 ) | tee gpurun_out/cn_golden.log
