#!/bin/bash
# Round-2 (session 3) capture set (under gpurun, 1 GPU).  usage: run_r2b_captures.sh TAG [quick]
# quick: ncu wave summaries only (stall reasons + pipe utilisation included); else also the bench lines.
tag=${1:-r02b}; mode=$2
mkdir -p gpurun_out
for sc in cbox room; do
  if [ $sc = cbox ]; then args="--spp 16 --reps 2"; wh="--width 1024 --height 1024"; else args="--scene room --res 1920 1080 --spp 8 --reps 2"; wh="--width 1920 --height 1080"; fi
  ncu --set full --import-source on --clock-control none -k regex:aq_k_ -s 18 -c 17 -f -o gpurun_out/${tag}_prof_$sc \
      python tools/quick_bench.py $args > gpurun_out/${tag}_ncu_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_prof_$sc.ncu-rep gpurun_out/${tag}_${sc}_ncu_full.json > /dev/null 2>> gpurun_out/${tag}_ncu_$sc.log
  python tools/ncu_wave_summary.py gpurun_out/${tag}_prof_$sc.ncu-rep --scene $sc $wh --pool 16777216 \
      --source ${tag}_${sc}_ncu_full.json --out gpurun_out/${tag}_ncu_summary_$sc.json > /dev/null 2>> gpurun_out/${tag}_ncu_$sc.log
  # source-line attribution of the depth-1 closest-hit launch and the depth-1 shade launch
  # depth-1 launches: the first launch of the DYN instantiations, the second shade launch
  AQ_LINE_KERNEL="aq_k_trace<(int)3, (bool)0, (int)6>" python tools/ncu_line_attrib.py gpurun_out/${tag}_prof_$sc.ncu-rep aq_k_trace _Z10aq_k_traceILi3ELb0ELi6 0 aq_kernels.cuh > gpurun_out/${tag}_lines_closest_d1_$sc.txt 2>> gpurun_out/${tag}_ncu_$sc.log
  AQ_LINE_KERNEL="aq_k_trace<(int)1, (bool)0, (int)6>" python tools/ncu_line_attrib.py gpurun_out/${tag}_prof_$sc.ncu-rep aq_k_trace _Z10aq_k_traceILi1ELb0ELi6 0 aq_kernels.cuh > gpurun_out/${tag}_lines_shadow_d1_$sc.txt 2>> gpurun_out/${tag}_ncu_$sc.log
  python tools/ncu_line_attrib.py gpurun_out/${tag}_prof_$sc.ncu-rep aq_k_shade _Z10aq_k_shadeILb0ELb0 1 aq_core.h > gpurun_out/${tag}_lines_shade_d1_$sc.txt 2>> gpurun_out/${tag}_ncu_$sc.log
done
rm -f gpurun_out/*.ncu-rep
if [ "$mode" != quick ]; then
  ( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
  python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c2_ref.json 2>> gpurun_out/${tag}_bench.err
  python bench.py --scene room --width 1920 --height 1080 --spp 256 --cpu-spp 1 --strong-spp 0 > gpurun_out/${tag}_bench_c3.json 2>> gpurun_out/${tag}_bench.err
  ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
      python bench.py --steps 1 --warmup 1 --spp 128 --no-cpu-baseline --strong-spp 0 > gpurun_out/${tag}_ncu_launches.log 2>&1
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
  python tools/bench_nrc.py --tensor > gpurun_out/${tag}_nrc_room_1080p.json 2>> gpurun_out/${tag}_bench.err
  python tools/bench_soup.py > gpurun_out/${tag}_soup_c4.json 2>> gpurun_out/${tag}_bench.err
  TAG=$tag bash tools/sanitize.sh > gpurun_out/${tag}_sanitize.log 2>&1
  tail -n 4 gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_bench.err gpurun_out/${tag}_smoke.log gpurun_out/${tag}_sanitize_summary.txt
fi
du -sh gpurun_out
