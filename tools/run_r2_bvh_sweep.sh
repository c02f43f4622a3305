#!/bin/bash
# builder knobs vs render time (the triangle loop runs at 5-10 of 32 lanes, the node visit at 23-24:
# is a tree with fewer triangle tests and more node visits faster?)
mkdir -p gpurun_out
for leaf in 3 2 1; do for ct in 0.3 0.6 1.0 2.0; do
  echo "== leaf=$leaf ct=$ct"
  AQUA_ACCEL_BUILDER=host AQUA_BVH_LEAF=$leaf AQUA_BVH_CT=$ct python tools/stage_bench.py --reps 3 2>/dev/null | grep prof=4 | cut -c1-200
  AQUA_ACCEL_BUILDER=host AQUA_BVH_LEAF=$leaf AQUA_BVH_CT=$ct python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 --reps 3 2>/dev/null | grep prof=4 | cut -c1-200
done; done > gpurun_out/r02_bvh_knob_sweep.log 2>&1
cat gpurun_out/r02_bvh_knob_sweep.log
