#!/bin/bash
mkdir -p gpurun_out
python tools/halo_timeline.py > gpurun_out/timeline.log 2>&1
for f in 1 2 3 0; do B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick7.log 2>&1; done
B200SD_FUSED=3 B200SD_HALO_MIN_HW=4096 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick7.log 2>&1
B200SD_STAGED=0 B200SD_FUSED=0 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick7.log 2>&1
timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_capi_gpu.py tests/test_factory_gpu.py tests/test_ops_gpu.py -q -m gpu > gpurun_out/tests7.log 2>&1; echo "rc=$?" >> gpurun_out/tests7.log
timeout 1500 python -m pytest tests/test_unet_gpu.py -q > gpurun_out/unet7.log 2>&1; echo "rc=$?" >> gpurun_out/unet7.log
cat gpurun_out/timeline.log | grep -v "chunk 1[0-9]\|chunk  [3-9]"; grep quick gpurun_out/quick7.log; tail -25 gpurun_out/tests7.log; tail -6 gpurun_out/unet7.log
