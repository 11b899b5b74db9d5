#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/quick9.log
for r in 1 0; do for f in ln 1 2; do B200SD_ROTATE=$r B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick9.log 2>&1; done; done
B200SD_ROTATE=1 timeout 300 python tools/halo_timeline.py > gpurun_out/timeline9_rot1.log 2>&1
B200SD_ROTATE=0 timeout 300 python tools/halo_timeline.py > gpurun_out/timeline9_rot0.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/tests9.log 2>&1; echo "rc=$?" >> gpurun_out/tests9.log
grep quick gpurun_out/quick9.log; grep -h "us" gpurun_out/timeline9_rot1.log | tail -12; echo ---; grep -h "us" gpurun_out/timeline9_rot0.log | tail -12; tail -15 gpurun_out/tests9.log | cut -c1-200
