#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -3 ) | tee gpurun_out/tests.log
for cfg in "0 2048" "1 2048" "0 100000" "1 100000" "0 2048"; do
  set -- $cfg
  echo "fold=$1 t512_min=$2"
  ( B200SD_GN_FOLD=$1 B200SD_GN_T512_MIN=$2 timeout 400 python bench.py --steps 40 --warmup 3 --no-batched 2>&1 | tail -1 ) | tee gpurun_out/bench_$1_$2.log | cut -c1-120
done
B200SD_GN_FOLD=1 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k group_norm 2>&1 | tail -2
( timeout 600 python tools/bench_configs.py 2>&1 | tail -2 ) | tee gpurun_out/configs.log
