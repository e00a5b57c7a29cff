#!/bin/bash
# End-of-round visit: the GPU tests, then the evidence set (tools/r2_profile.sh).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r02d}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log
tools/r2_profile.sh $TAG
