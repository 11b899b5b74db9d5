#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pdl_ab.log
for r in 1 2; do for p in 0 1; do B200SD_PDL=$p timeout 300 python bench.py --quick --steps 40 2>&1 | grep quick >> gpurun_out/pdl_ab.log; done; done
B200SD_PDL=1 timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/tests_pdl.log 2>&1; echo "rc=$?" >> gpurun_out/tests_pdl.log
cat gpurun_out/pdl_ab.log; tail -5 gpurun_out/tests_pdl.log | cut -c1-250
