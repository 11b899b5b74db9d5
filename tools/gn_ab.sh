#!/bin/bash
mkdir -p gpurun_out
for t in 256 512; do B200SD_GN_THREADS=$t timeout 300 python bench.py --quick --steps 40 2>&1 | grep quick; B200SD_GN_THREADS=$t timeout 300 python tools/halo_timeline.py 2>&1 | grep "cluster GN" | cut -c1-18,100-125; done > gpurun_out/gn_threads.log 2>&1
cat gpurun_out/gn_threads.log
