#!/bin/bash
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) | tee gpurun_out/tests.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/smoke.log
( timeout 500 python bench.py --steps 40 --warmup 3 2>&1 | tail -2 ) | tee gpurun_out/bench.log | cut -c1-300
