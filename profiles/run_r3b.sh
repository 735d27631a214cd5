#!/bin/bash
# per-instruction warp-stall sampling of the middle LDE pass (ntt_pass_kernel launch #5 of one LDE)
NCU="ncu --clock-control none --profile-from-start off"
$NCU --set full --import-source on -k regex:ntt_pass -s 4 -c 1 -f -o /tmp/r3b python profiles/prof_run.py lde > gpurun_out/r3b_prof.log 2>&1
ncu -i /tmp/r3b.ncu-rep --page source --csv > gpurun_out/r3b_source.csv 2>> gpurun_out/r3b_prof.log
wc -l gpurun_out/r3b_source.csv; head -c 1500 gpurun_out/r3b_source.csv
