#!/bin/bash
# Round 2: --set full of the kernels added this round (argument builders, lookup sort, point decompression)
NCU="ncu --clock-control none --profile-from-start off"
$NCU --set full --import-source on -k 'regex:expr_eval|quotient_div|ls_|decompress' -c 70 -f -o /tmp/r2x python profiles/prof_new_kernels.py > gpurun_out/r2x_prof.log 2>&1
python profiles/ncu_summary.py /tmp/r2x.ncu-rep > gpurun_out/r2x_full_new_kernels.txt 2>&1
tail -3 gpurun_out/r2x_prof.log
