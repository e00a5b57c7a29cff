#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v30}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log | head -5
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep rep1 | tee $OUT/${TAG}_ab.txt
