#!/bin/bash
mkdir -p gpurun_out
( B200SD_DEBUG_SYNC=1 timeout 60 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -s -k "test_conv3x3" 2>&1 | tail -15 ) | tee gpurun_out/dbg_conv.log
timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:gn_stats -s 2 -c 2 -o gpurun_out/gn_r1c python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_gn.log 2>&1
ls -la gpurun_out/gn_r1c.ncu-rep
