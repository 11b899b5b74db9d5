#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/tests16.log 2>&1; echo "rc=$?" >> gpurun_out/tests16.log
tail -15 gpurun_out/tests16.log | cut -c1-300
