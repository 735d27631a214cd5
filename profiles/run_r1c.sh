#!/bin/bash
# Round-1 (final kernels) ncu captures, one B200 under gpurun.  Same recipe as run_r1b.sh; outputs gpurun_out/r1c_*.
NCU="ncu --clock-control none --profile-from-start off"
for m in lde ntt24 "msm 20" "msm 20 20" "lpc 0"; do
  tag=$(echo $m | tr ' ' '_')
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1c_launches_${tag}.csv python profiles/prof_run.py $m > gpurun_out/r1c_prof_${tag}.log 2>&1
done
full() {  # name, kernel regex, count, workload...
  name=$1; k=$2; c=$3; shift 3
  $NCU --set full --import-source on -k regex:$k -c $c -f -o /tmp/r1c_$name python profiles/prof_run.py "$@" >> gpurun_out/r1c_prof_full.log 2>&1
  python profiles/ncu_summary.py /tmp/r1c_$name.ncu-rep > gpurun_out/r1c_full_$name.txt 2>&1
}
full ntt24 ntt_pass 3 ntt24
full lde ntt_pass 6 lde
full msm_acc msm_accumulate 1 msm 20 20
full msm_red msm_reduce 7 msm 20 20
full leaf leaf_hash 1 lpc 0
ls -la gpurun_out/
