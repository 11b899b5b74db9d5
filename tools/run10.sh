#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention" > gpurun_out/tests10.log 2>&1; echo "rc=$?" >> gpurun_out/tests10.log
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
timeout 300 python bench.py --quick --steps 40 > gpurun_out/quick10.log 2>&1
B200SD_ATTN_STREAMK=0 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick10.log 2>&1
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_capi_gpu.py -q -m gpu > gpurun_out/tests10b.log 2>&1; echo "rc=$?" >> gpurun_out/tests10b.log
tail -12 gpurun_out/tests10.log | cut -c1-250; cat gpurun_out/attn_bench.log | tail -8; grep quick gpurun_out/quick10.log; tail -5 gpurun_out/tests10b.log | cut -c1-250
