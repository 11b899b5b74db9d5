#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -q -m gpu -k "shortcut" > gpurun_out/tests_sc.log 2>&1; echo "rc=$?" >> gpurun_out/tests_sc.log
for v in 1 0 1 0; do B200SD_FOLD_SC=$v timeout 300 python bench.py --quick --steps 40 2>&1 | grep quick; done > gpurun_out/sc_ab.log
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_capi_gpu.py tests/test_factory_gpu.py -q -m gpu > gpurun_out/tests_sc2.log 2>&1; echo "rc=$?" >> gpurun_out/tests_sc2.log
tail -6 gpurun_out/tests_sc.log | cut -c1-300; cat gpurun_out/sc_ab.log; tail -8 gpurun_out/tests_sc2.log | cut -c1-300
