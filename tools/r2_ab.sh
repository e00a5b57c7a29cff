#!/bin/bash
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | tee gpurun_out/${1:-ab}.txt
