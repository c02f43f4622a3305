#!/bin/bash
# ncu --set full of the depth-1 closest-hit + shadow launches of a cbox / room wave for several library variants
mkdir -p gpurun_out
for v in "$@"; do
  lib=libaqua_cuda.so; [ "$v" != base ] && lib=libaqua_cuda_$v.so
  for sc in cbox room; do
    if [ $sc = cbox ]; then args="--spp 16 --reps 2"; else args="--scene room --res 1920 1080 --spp 8 --reps 2"; fi
    AQUA_CUDA_LIB=$lib ncu --set full --clock-control none -k regex:aq_k_trace -s 12 -c 2 -f -o gpurun_out/ab_${v}_$sc python tools/quick_bench.py $args > /dev/null 2>&1
    echo "== $v $sc"; python tools/ncu_summary.py gpurun_out/ab_${v}_$sc.ncu-rep | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.items() if k in ('duration_us','warp_inst','threads_per_inst','issue_active_pct','achieved_occupancy_pct','regs','l1_hit_pct','smem_bank_conflicts','dram_gbs')})"
    ncu -i gpurun_out/ab_${v}_$sc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if 'issue_stalled' in h and 'ratio' in h or 'inst_executed_op_shared' in h or 'icc' in h.lower() or 'no_instruction' in h]
for r in rows[2:]:
    print({hdr[i].replace('smsp__average_warp_latency_issue_stalled_','').replace('smsp__average_warps_issue_stalled_','st_'): r[i] for i in keep if r[i] not in ('0','')})" | cut -c1-1500
    rm -f gpurun_out/ab_${v}_$sc.ncu-rep
  done
done
