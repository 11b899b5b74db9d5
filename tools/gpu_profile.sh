#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/smoke.log
( timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -15 ) | tee gpurun_out/unet_tests.log
( timeout 600 python tools/profile_unet.py --kernels --shapes 2>&1 | tail -80 ) | tee gpurun_out/shapes.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 529 -c 529 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_unet.py --forwards 2 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
