#!/bin/bash
# Round 2: --set full of every kernel of one tabled MSM at 2^20 (c = 20) and 2^16 (c = 18): where the reduce tail goes.
NCU="ncu --clock-control none --profile-from-start off"
for cfg in "20 20" "16 18"; do
  set -- $cfg
  $NCU --set full --import-source on -f -o /tmp/r2l_msm_$1 python profiles/prof_run.py msm $1 $2 > gpurun_out/r2l_prof_$1.log 2>&1
  python profiles/ncu_summary.py /tmp/r2l_msm_$1.ncu-rep > gpurun_out/r2l_full_msm_$1.txt 2>&1
done
ls -la gpurun_out | tail -5
