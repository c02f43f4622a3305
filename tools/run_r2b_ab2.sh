#!/bin/bash
# round 2 (session 3) A/B with run-time switches: usage run_r2b_ab2.sh TAG "lib|ENV=.. ENV=.." ...
tag=$1; shift
mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then ( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1; else : > gpurun_out/${tag}_pytest_gpu.log; fi
for rep in 1 2; do
for v in "$@"; do
  lib=${v%%|*}; envs=${v#*|}
  for sc in "" "--scene room --res 1920 1080 --spp 8"; do
    env AQUA_CUDA_LIB=$lib $envs python tools/stage_bench.py $sc 2>/dev/null | grep "prof=4" | sed "s/^/[$envs] /"
  done
done
done > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_pytest_gpu.log; sort gpurun_out/${tag}_ab.log
