#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "group_norm" 2>&1 | tail -5 ) | tee gpurun_out/gn.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/smoke.log
( timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8 ) | tee gpurun_out/unet_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 4 -c 6 -o gpurun_out/gemm_r1 python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 2 -o gpurun_out/attn_r1 python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gn_ -s 2 -c 4 -o gpurun_out/gn_r1 python tools/profile_unet.py --forwards 1 > gpurun_out/ncu_gn.log 2>&1
ls -la gpurun_out/
