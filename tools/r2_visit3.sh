#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v3}
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fine|k_heavy" -s 10 -c 2 -f -o $OUT/${TAG}_fine \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --frame-events > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}_fine.ncu-rep --page raw --csv > $OUT/${TAG}_fine_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_fine.ncu-rep --page source --csv > $OUT/${TAG}_fine_source.csv 2>/dev/null
ls -la $OUT/${TAG}_*
