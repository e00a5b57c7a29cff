#!/bin/bash
# two GPUs: the group tests and the torchrun bench
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
MODES=balanced tools/scale_bench.sh 2 2>&1 | tee gpurun_out/v35_scale2.txt
