#!/bin/bash
# parity tests on an A/B library: usage ab_verify2.sh lib.so
echo "== parity tests on $1"
AQUA_CUDA_LIB=$1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q 2>&1 | tail -4
