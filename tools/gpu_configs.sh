#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:umma_gemm --launch-skip 8 --launch-count 4 -o gpurun_out/pair_r1 python tools/profile_pair_gemm.py > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
