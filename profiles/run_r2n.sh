#!/bin/bash
# Round 2: --set full of the six passes of one LDE (16 x 2^20 -> 2^23 Pallas Fq) and the three of a coset NTT 2^24 after the
# block-twiddle rewrite (no inter-pass tables).  usage: bash profiles/run_r2n.sh TAG
TAG=${1:-r2n}
NCU="ncu --clock-control none --profile-from-start off"
$NCU --set full --import-source on -k regex:ntt_ -c 7 -f -o /tmp/${TAG}_lde python profiles/prof_run.py lde > gpurun_out/${TAG}_prof.log 2>&1
python profiles/ncu_summary.py /tmp/${TAG}_lde.ncu-rep > gpurun_out/${TAG}_full_lde.txt 2>&1
$NCU --set full --import-source on -k regex:ntt_pass -c 3 -f -o /tmp/${TAG}_ntt24 python profiles/prof_run.py ntt24 >> gpurun_out/${TAG}_prof.log 2>&1
python profiles/ncu_summary.py /tmp/${TAG}_ntt24.ncu-rep > gpurun_out/${TAG}_full_ntt24.txt 2>&1
ls -la gpurun_out | tail -4
