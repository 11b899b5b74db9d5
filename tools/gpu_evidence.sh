#!/bin/bash
# evidence refresh on the final build: ncu launch list / DRAM bytes / --set full rows of ONE bench.py iteration
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum --cache-control none --csv --log-file gpurun_out/launches_r2_final.csv \
  python bench.py --profile-step > gpurun_out/ncu_l1.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
  --log-file gpurun_out/launches_r2_final_dram.csv python bench.py --profile-step > gpurun_out/ncu_l2.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:attention_kernel -c 2 -o gpurun_out/attn_r2_final -f \
  python bench.py --profile-step > gpurun_out/ncu_attn.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:gn_cluster -c 3 -o gpurun_out/gn_r2_final -f \
  python bench.py --profile-step > gpurun_out/ncu_gn.log 2>&1
timeout 1200 $NCU --set full --import-source on -k regex:umma_gemm -s 2 -c 20 -o gpurun_out/gemm_r2_final -f \
  python bench.py --profile-step > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_final*.csv; tail -2 gpurun_out/ncu_l1.log
