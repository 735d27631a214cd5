#!/bin/bash
# Round 2: --set full of the lookup-sort and point-decompression kernels
NCU="ncu --clock-control none --profile-from-start off"
$NCU --set full --import-source on -k 'regex:^(ls_|g1_decompress|g2_decompress)' -c 12 -f -o /tmp/r2y python profiles/prof_new_kernels.py keys > gpurun_out/r2y_prof.log 2>&1
python profiles/ncu_summary.py /tmp/r2y.ncu-rep > gpurun_out/r2y_full_lookup_decompress.txt 2>&1
tail -2 gpurun_out/r2y_prof.log
