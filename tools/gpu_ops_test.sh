#!/bin/bash
# Runs the op-level GPU parity tests group by group (a crashing kernel must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in linear conv3x3 group_norm layer_norm attention small_kernels cfg_scheduler; do
  echo "=== $grp ===" | tee -a gpurun_out/ops.log
  timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$grp" 2>&1 | tail -25 | tee -a gpurun_out/ops.log
done
