#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "pairs" 2>&1 | tail -3 ) | tee gpurun_out/tests_pairs.log
( B200SD_PDL=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -5 ) | tee gpurun_out/tests_pdl.log
for v in 0 1 0 1; do
  echo "pdl=$v"
  ( B200SD_PDL=$v timeout 400 python bench.py --steps 40 --warmup 3 --no-batched 2>&1 | tail -1 ) | tee gpurun_out/bench_pdl_$v.log | cut -c1-120
done
