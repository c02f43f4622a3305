#!/bin/bash
# Session-2 GPU call: parity of the full-Principled shade instantiations + wavefront pool sweep.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/s2_pytest.log
{
for pool in 0 16777216 33554432; do
  python tools/stage_bench.py --pool $pool
  python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 --pool $pool
done
python tools/stage_bench.py --full-bsdf
python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 --full-bsdf
} 2>&1 | tee gpurun_out/s2_stage.log
