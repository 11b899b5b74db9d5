#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_warm_r1.csv python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:gn_fused -c 3 -o gpurun_out/gnf_r1 python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_gnf.log 2>&1
