#!/bin/bash
# C4 soup: balanced triangle phase on / off in the aq_intersect kernels
mkdir -p gpurun_out
for e in 0 1; do AQUA_TRI_DYN=$e python tools/bench_soup.py --brief --rays 33554432 --check-bvh 262144 2>/dev/null | sed "s/^/[AQUA_TRI_DYN=$e] /"; done > gpurun_out/r02b_soup_ab.log 2>&1
cat gpurun_out/r02b_soup_ab.log
