#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in tl tl3; do echo "== $v"; PM_LIB=$PWD/piet-metal_b200/variants/libpm_$v.so python tools/grid_timeline.py 8192 2>&1 | tail -26; done | tee $OUT/v39_timeline.txt
