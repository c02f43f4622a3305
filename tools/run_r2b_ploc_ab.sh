#!/bin/bash
# device builder: PLOC vs radix tree — parity on the device-builder tests, room on the device-built tree, C4 soup
mkdir -p gpurun_out
( echo "== device-builder parity tests with AQUA_DEVICE_TREE=ploc"; AQUA_DEVICE_TREE=ploc timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device or lbvh or hybrid or soup" 2>&1 | tail -3 ) > gpurun_out/r02c_ploc_verify.log 2>&1
cat gpurun_out/r02c_ploc_verify.log
for t in lbvh ploc; do
  for rep in 1 2; do AQUA_DEVICE_TREE=$t AQUA_ACCEL_BUILDER=device python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 2>/dev/null | grep "prof=4" | sed "s/^/[tree=$t] /"; done
  AQUA_DEVICE_TREE=$t AQUA_ACCEL_BUILDER=device AQ_BUILD_VERBOSE=1 python tools/quick_bench.py --scene room --res 1920 1080 --spp 8 --reps 1 2>&1 | grep -E "upload|build" | sed "s/^/[tree=$t] /"
  AQUA_DEVICE_TREE=$t python tools/bench_soup.py --brief --rays 33554432 --check-bvh 262144 2>/dev/null | sed "s/^/[tree=$t] /"
done | tee gpurun_out/r02c_ploc_ab.log
