#!/bin/bash
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -15
