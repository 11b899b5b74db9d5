#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -q > gpurun_out/fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/fused_tests.log
timeout 600 python -m pytest tests/test_ops_gpu.py -q > gpurun_out/ops_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ops_tests.log
timeout 1500 python -m pytest tests/test_unet_gpu.py -q > gpurun_out/unet_tests.log 2>&1; echo "rc=$?" >> gpurun_out/unet_tests.log
timeout 600 python tools/profile_unet.py --forwards 2 --shapes > gpurun_out/shapes_fused.log 2>&1
B200SD_FUSED=0 timeout 600 python tools/profile_unet.py --forwards 2 --shapes > gpurun_out/shapes_unfused.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_r2_fused_v1.csv python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:halo_conv -s 2 -c 6 \
  -o gpurun_out/halo_v1 -f python tools/profile_unet.py --forwards 2 --capture-last > gpurun_out/ncu_halo.log 2>&1
tail -12 gpurun_out/fused_tests.log; tail -4 gpurun_out/ops_tests.log; tail -12 gpurun_out/unet_tests.log
python tools/summarize_launches.py gpurun_out/launches_r2_fused_v1.csv | head -24
for f in 1 0; do B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick3.log 2>&1; done
grep quick gpurun_out/quick3.log
head -45 gpurun_out/shapes_fused.log
