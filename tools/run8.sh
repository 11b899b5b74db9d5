#!/bin/bash
mkdir -p gpurun_out
for f in ln 0 1 2; do B200SD_FUSED=$f timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick8.log 2>&1; done
B200SD_PDL=1 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick8.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/tests8.log 2>&1; echo "rc=$?" >> gpurun_out/tests8.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
grep quick gpurun_out/quick8.log; tail -12 gpurun_out/tests8.log | cut -c1-200; tail -3 gpurun_out/bench_full.log | cut -c1-3000; tail -2 gpurun_out/bench_ref.log | cut -c1-600; tail -2 gpurun_out/smoke.log
