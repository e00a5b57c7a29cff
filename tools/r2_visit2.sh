#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $OUT/v2_pytest.log 2>&1; tail -15 $OUT/v2_pytest.log
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 3 > $OUT/v2_bench.json 2> $OUT/v2_bench.err; cat $OUT/v2_bench.json; tail -3 $OUT/v2_bench.err
