#!/bin/bash
mkdir -p gpurun_out
B200SD_PDL=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attention_kernel -c 3 -o gpurun_out/attn_r1b python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_attn.log 2>&1
B200SD_PDL=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:umma_gemm_kernel -s 4 -c 5 -o gpurun_out/gemm_r1b python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
