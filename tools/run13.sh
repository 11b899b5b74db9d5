#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -q -m gpu -k "tma or plain" > gpurun_out/tests13.log 2>&1; echo "rc=$?" >> gpurun_out/tests13.log
timeout 600 python tools/halo_tma_ab.py > gpurun_out/halo_tma_ab.log 2>&1
tail -15 gpurun_out/tests13.log | cut -c1-300; cat gpurun_out/halo_tma_ab.log | tail -14
