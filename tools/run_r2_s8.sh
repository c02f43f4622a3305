#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nrc.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/s8_pytest_nrc.log 2>&1
( timeout 600 python tools/bench_nrc.py --tensor 2>&1 | tail -1 ) > gpurun_out/s8_bench_nrc_room.json 2>&1
( timeout 600 python tools/bench_nrc.py --tensor --scene cbox --res 1024 1024 2>&1 | tail -1 ) > gpurun_out/s8_bench_nrc_cbox.json 2>&1
for v in base tol; do
  lib=libaqua_cuda.so; [ "$v" != base ] && lib=libaqua_cuda_$v.so
  AQUA_CUDA_LIB=$lib python tools/stage_bench.py 2>/dev/null | grep prof=4
  AQUA_CUDA_LIB=$lib python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null | grep prof=4
done > gpurun_out/s8_ab_tol.log 2>&1
cat gpurun_out/s8_*
