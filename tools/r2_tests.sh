#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-t}
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/${TAG}_pytest.log 2>&1; tail -25 $OUT/${TAG}_pytest.log
