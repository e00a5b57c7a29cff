#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v25}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | tee $OUT/${TAG}_ab.txt
echo "== PM_L2_PERSIST=1"; PM_L2_PERSIST=1 BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep "b200.so" | tee -a $OUT/${TAG}_ab.txt
echo "== strips 8192"; SPECS="2:0 8:0 8:3" tools/strip_study.sh 8192 2>&1 | tee $OUT/${TAG}_strips.txt
