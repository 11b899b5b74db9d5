#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/tests.log
( timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee gpurun_out/smoke.log
