#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v6}
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep rep2 | tee $OUT/${TAG}_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fine|k_heavy|k_seg|k_row" -s 12 -c 4 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --frame-events > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof.ncu-rep --page source --csv > $OUT/${TAG}_source.csv 2>/dev/null
ls -la $OUT/${TAG}_*
