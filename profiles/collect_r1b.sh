#!/bin/bash
# Copies the summaries of run_r1b.sh from gpurun_out/ (scratch) to the tracked profiles/ directory.
cd "$(dirname "$0")/.."
for f in gpurun_out/r1b_launches_*.csv; do grep -v "^==" "$f" > profiles/$(basename "$f"); done
for f in gpurun_out/r1b_full_*.txt; do cp "$f" profiles/$(basename "$f"); done
