#!/bin/bash
# round 2 (session 3) A/B: usage run_r2b_ab.sh TAG VARIANT...   (base = libaqua_cuda.so, X = libaqua_cuda_X.so)
tag=$1; shift
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
extra=$(cat tools/ab_variants.txt 2>/dev/null)
for rep in 1 2; do
for v in "$@" $extra; do
  lib=libaqua_cuda.so; [ "$v" != base ] && lib=libaqua_cuda_$v.so
  export AQUA_CUDA_LIB=$lib
  python tools/stage_bench.py 2>/dev/null | grep "prof=4"
  python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null | grep "prof=4"
done
done > gpurun_out/${tag}_ab.log 2>&1
for v in $(cat tools/ab_verify.txt 2>/dev/null); do
  echo "== parity tests on libaqua_cuda_$v.so"
  AQUA_CUDA_LIB=libaqua_cuda_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q 2>&1 | tail -3
done > gpurun_out/${tag}_verify.log 2>&1
cat gpurun_out/${tag}_verify.log gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_ab.log
