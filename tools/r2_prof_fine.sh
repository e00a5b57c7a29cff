#!/bin/bash
# One full ncu capture (with source) of k_fine on the headline frame.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-pf}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fine" -s 10 -c 1 -f -o $OUT/${TAG}_fine \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}_fine.ncu-rep --page raw --csv > $OUT/${TAG}_fine_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_fine.ncu-rep --page source --csv > $OUT/${TAG}_fine_source.csv 2>/dev/null
ls -la $OUT/${TAG}_*
