#!/bin/bash
# source-line attribution of the closest-hit kernel at depth 1 (cbox C2 wave, room 1080p wave); the
# .ncu-rep is processed here and deleted
mkdir -p gpurun_out
for sc in cbox room; do
  if [ $sc = cbox ]; then args="--spp 16 --reps 2"; else args="--scene room --res 1920 1080 --spp 8 --reps 2"; fi
  # launches of the 2nd render: 18 raygen, 19 closest d0, 20 shade, 21 shadow, 22 closest d1
  ncu --set full --clock-control none --import-source on -k regex:aq_k_ -s 22 -c 1 -f -o gpurun_out/lines_$sc python tools/quick_bench.py $args > gpurun_out/lines_$sc.log 2>&1
  python tools/ncu_line_attrib.py gpurun_out/lines_$sc.ncu-rep aq_k_trace _Z10aq_k_traceILi3ELb0 0 aq_kernels.cuh > gpurun_out/r02_lines_closest_d1_${sc}_by_kernel_line.txt 2>&1
  python tools/ncu_line_attrib.py gpurun_out/lines_$sc.ncu-rep aq_k_trace _Z10aq_k_traceILi3ELb0 0 aq_bvh.h > gpurun_out/r02_lines_closest_d1_${sc}_by_bvh_line.txt 2>&1
  rm -f gpurun_out/lines_$sc.ncu-rep
done
head -50 gpurun_out/r02_lines_closest_d1_cbox_by_bvh_line.txt
