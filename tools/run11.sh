#!/bin/bash
# round-2 evidence run: ncu launch list / DRAM bytes / --set full rows of ONE bench.py iteration, plus PDL and split-K A/B
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
rm -f gpurun_out/quick11.log
timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick11.log 2>&1
B200SD_CLUSTER_SPLITK=0 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick11.log 2>&1
B200SD_PDL=1 timeout 300 python bench.py --quick --steps 40 >> gpurun_out/quick11.log 2>&1
B200SD_PDL=1 timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests11_pdl.log 2>&1; echo "rc=$?" >> gpurun_out/tests11_pdl.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_final*.csv
grep quick gpurun_out/quick11.log; tail -5 gpurun_out/tests11_pdl.log | cut -c1-200; tail -2 gpurun_out/ncu_l1.log
