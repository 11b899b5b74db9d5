#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention" > gpurun_out/tests15.log 2>&1; echo "rc=$?" >> gpurun_out/tests15.log
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench15.log 2>&1
timeout 300 python bench.py --quick --steps 40 > gpurun_out/quick15.log 2>&1
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_factory_gpu.py tests/test_capi_gpu.py -q -m gpu > gpurun_out/tests15b.log 2>&1; echo "rc=$?" >> gpurun_out/tests15b.log
tail -12 gpurun_out/tests15.log | cut -c1-250; cat gpurun_out/attn_bench15.log | tail -8; grep quick gpurun_out/quick15.log; tail -12 gpurun_out/tests15b.log | cut -c1-250
