#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v26}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log
for spec in "" "8:0" "8:3"; do echo "== timeline 8192 $spec"; PM_LIB=$PWD/piet-metal_b200/variants/libpm_tl.so python tools/grid_timeline.py 8192 $spec 2>&1 | tail -12; done | tee $OUT/${TAG}_timeline.txt
echo "== strips 8192"; SPECS="1:0 8:0 8:3" tools/strip_study.sh 8192 2>&1 | tee $OUT/${TAG}_strips.txt
