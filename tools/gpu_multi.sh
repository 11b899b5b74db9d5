#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 --no-batched 2>&1 | tail -3 ) | tee gpurun_out/bench_n2.log | cut -c1-400
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 ) | tee gpurun_out/bench_ref_n2.log | cut -c1-300
