#!/bin/bash
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | grep rep1 | tee gpurun_out/v43_ab.txt
