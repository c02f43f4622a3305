#!/bin/bash
# A/B of kernel variants built with `build.py --variant NAME DEFS` (selected by AQUA_CUDA_LIB).
# usage: tools/run_ab.sh NAME [NAME...]   ("base" = libaqua_cuda.so)
mkdir -p gpurun_out
{
for v in "$@"; do
  lib=libaqua_cuda.so; [ "$v" != base ] && lib=libaqua_cuda_$v.so
  export AQUA_CUDA_LIB=$lib
  if [ "$v" != base ] && [ -z "$AB_SKIP_TESTS" ]; then python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1; fi
  python tools/stage_bench.py 2>/dev/null
  python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null
done
} 2>&1 | tee gpurun_out/ab_$1_$2.log
