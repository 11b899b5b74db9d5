#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "group_norm" > gpurun_out/tests17.log 2>&1; echo "rc=$?" >> gpurun_out/tests17.log
timeout 300 python bench.py --quick --steps 40 > gpurun_out/quick17.log 2>&1
timeout 400 python tools/halo_timeline.py 2>&1 | grep "cluster GN" > gpurun_out/gn17.log
tail -4 gpurun_out/tests17.log | cut -c1-250; grep quick gpurun_out/quick17.log; cat gpurun_out/gn17.log | cut -c100-200
