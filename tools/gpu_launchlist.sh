#!/bin/bash
# Launch list of one warm UNet forward (durations; warm L2), the same list with DRAM bytes (ncu's default cache
# flush between kernels), and a full-section capture of the cluster GroupNorm kernel.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_warm_r1.csv python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_dram_r1.csv python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_dram.log 2>&1
tail -2 gpurun_out/ncu_dram.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:gn_cluster -c 4 -o gpurun_out/gnc_r1 python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_gnc.log 2>&1
tail -2 gpurun_out/ncu_gnc.log
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) | tee gpurun_out/tests.log
