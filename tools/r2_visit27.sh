#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v27}
for spec in "" "8:0" "8:3"; do echo "== timeline 8192 $spec"; PM_LIB=$PWD/piet-metal_b200/variants/libpm_tl.so python tools/grid_timeline.py 8192 $spec 2>&1 | tail -12; done | tee $OUT/${TAG}_timeline.txt
