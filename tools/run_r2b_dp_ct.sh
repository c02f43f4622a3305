#!/bin/bash
# device DP collapse: triangle / node cost ratio sweep (room on the device-built tree, C4 soup)
lib=${1:-libaqua_cuda.so}
mkdir -p gpurun_out
for ct in 0.3 0.6 1.0 1.5 2.5; do
  AQUA_CUDA_LIB=$lib AQUA_ACCEL_BUILDER=device AQUA_BVH_CT=$ct python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null | grep "prof=4" | sed "s/^/[ct=$ct] /"
  AQUA_CUDA_LIB=$lib AQUA_BVH_CT=$ct python tools/bench_soup.py --brief --rays 33554432 --check-bvh 262144 2>/dev/null | sed "s/^/[ct=$ct] /"
done | tee gpurun_out/r02b_dp_ct.log
