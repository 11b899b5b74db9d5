#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -12 ) | tee gpurun_out/tests.log
( timeout 400 python bench.py --steps 40 --warmup 3 2>&1 | tail -2 ) | tee gpurun_out/bench.log | cut -c1-400
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 ) | tee gpurun_out/bench_ref.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 458 -c 458 --csv --log-file gpurun_out/launches_warm_r1.csv python tools/profile_unet.py --forwards 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
