#!/bin/bash
# A/B of kernel variants built with `build.py --variant NAME DEFS`; usage: run_r2_ab.sh TAG NAME...
tag=$1; shift
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
for v in "$@"; do
  lib=libaqua_cuda.so; [ "$v" != base ] && lib=libaqua_cuda_$v.so
  export AQUA_CUDA_LIB=$lib
  python tools/stage_bench.py 2>/dev/null
  python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null
  [ -n "$SWEEP" ] && python tools/ctrl_sweep.py 2>/dev/null
done > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_ab.log
