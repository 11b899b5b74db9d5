#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/tests20.log 2>&1; echo "rc=$?" >> gpurun_out/tests20.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -4 gpurun_out/tests20.log | cut -c1-200; tail -3 gpurun_out/bench_full.log | cut -c1-3500; tail -2 gpurun_out/bench_ref.log | cut -c1-300; tail -2 gpurun_out/smoke.log
