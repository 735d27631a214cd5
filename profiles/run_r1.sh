#!/bin/bash
# Round-1 ncu captures (run under gpurun on one B200).  Outputs land in gpurun_out/.
set -x
NCU="ncu --clock-control none --profile-from-start off"
for m in lde ntt24 msm "lpc 0" "lpc 1"; do
  tag=$(echo $m | tr ' ' '_')
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}.csv python profiles/prof_run.py $m > gpurun_out/prof_${tag}.log 2>&1
done
$NCU --set full --import-source on -k regex:ntt_pass -c 6 -f -o gpurun_out/full_ntt24 python profiles/prof_run.py ntt24 >> gpurun_out/prof_full.log 2>&1
$NCU --set full --import-source on -k regex:msm_accumulate -c 1 -f -o gpurun_out/full_msm_acc python profiles/prof_run.py msm >> gpurun_out/prof_full.log 2>&1
$NCU --set full --import-source on -k regex:leaf_hash -c 1 -f -o gpurun_out/full_leaf_keccak python profiles/prof_run.py lpc 0 >> gpurun_out/prof_full.log 2>&1
ls -la gpurun_out/
