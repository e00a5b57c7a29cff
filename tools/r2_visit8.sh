#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v8}
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=6 ) > $OUT/${TAG}_pytest.log 2>&1; tail -14 $OUT/${TAG}_pytest.log
for lib in piet-metal_b200/variants/*.so; do echo "== $lib"; PM_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scene_matches or fuzz or deep or many_records or dense" 2>&1 | tail -2; done
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep rep2 | tee $OUT/${TAG}_ab.txt
