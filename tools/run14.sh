#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench_n2.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 ) > gpurun_out/bench_ref_n2.log
cut -c1-2500 gpurun_out/bench_n2.log; cut -c1-400 gpurun_out/bench_ref_n2.log
