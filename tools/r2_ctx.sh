#!/bin/bash
timeout 600 python -m pytest tests/test_context.py tests/test_gpu_flatten.py -m gpu -x -q 2>&1 | tail -15
