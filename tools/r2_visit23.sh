#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v23}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log | head -2
for c in 4 3 2 1; do echo "== PM_DEBUG_FINE_CTAS=$c"; PM_DEBUG_FINE_CTAS=$c BENCH_ARGS="--steps 100" tools/ab_bench.sh 2>&1 | grep rep1; done | tee $OUT/${TAG}_ctas.txt
