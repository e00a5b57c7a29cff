#!/bin/bash
# A/B of every library under piet-metal_b200/variants/ against the default one (kernel times and the frame)
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep rep1 | tee gpurun_out/${1:-ab}.txt
