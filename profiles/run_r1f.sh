#!/bin/bash
# Round-1 end-of-round ncu captures, one B200 under gpurun.  Same recipe as run_r1c.sh; outputs gpurun_out/r1f_*.
NCU="ncu --clock-control none --profile-from-start off"
for m in lde ntt24 "msm 20" "msm 20 20" "lpc 0" evalpm grind; do
  tag=$(echo $m | tr ' ' '_')
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1f_launches_${tag}.csv python profiles/prof_run.py $m > gpurun_out/r1f_prof_${tag}.log 2>&1
done
full() {  # name, kernel regex, count, workload...
  name=$1; k=$2; c=$3; shift 3
  $NCU --set full --import-source on -k regex:$k -c $c -f -o /tmp/r1f_$name python profiles/prof_run.py "$@" >> gpurun_out/r1f_prof_full.log 2>&1
  python profiles/ncu_summary.py /tmp/r1f_$name.ncu-rep > gpurun_out/r1f_full_$name.txt 2>&1
}
full lde ntt_pass 6 lde
full ntt24 ntt_pass 3 ntt24
full msm_acc msm_accumulate 1 msm 20 20
full evalpm poly_eval_segments 1 evalpm
full grind pow_grind 1 grind
full leaf leaf_hash 1 lpc 0
ls -la gpurun_out/
