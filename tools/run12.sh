#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention" > gpurun_out/tests12.log 2>&1; echo "rc=$?" >> gpurun_out/tests12.log
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench12.log 2>&1
timeout 300 python bench.py --quick --steps 40 > gpurun_out/quick12.log 2>&1
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_text_gpu.py -q -m gpu > gpurun_out/tests12b.log 2>&1; echo "rc=$?" >> gpurun_out/tests12b.log
tail -12 gpurun_out/tests12.log | cut -c1-250; cat gpurun_out/attn_bench12.log | tail -8; grep quick gpurun_out/quick12.log; tail -5 gpurun_out/tests12b.log | cut -c1-250
