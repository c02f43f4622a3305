#!/bin/bash
# Round-2 capture set (under gpurun, 1 GPU).  usage: run_r2_captures.sh TAG [quick]
tag=${1:-r02}; mode=$2
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
( python tools/stage_bench.py; python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 ) > gpurun_out/${tag}_stage.log 2>/dev/null
# ncu --set full of ONE full C2 wave (raygen + 5 x (closest, shade, shadow) + film = 17 launches), second render of the process
# (the .ncu-rep files are summarised HERE and deleted: gpurun brings back at most 64 MiB)
ncu --set full --clock-control none -k regex:aq_k_ -s 18 -c 17 -f -o gpurun_out/${tag}_prof_cbox \
    python tools/quick_bench.py --spp 16 --reps 2 > gpurun_out/${tag}_ncu_cbox.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_prof_cbox.ncu-rep gpurun_out/${tag}_cbox_ncu_full.json > /dev/null 2>> gpurun_out/${tag}_ncu_cbox.log
python tools/ncu_wave_summary.py gpurun_out/${tag}_prof_cbox.ncu-rep --scene cbox --width 1024 --height 1024 --pool 16777216 \
    --source ${tag}_cbox_ncu_full.json --out gpurun_out/${tag}_ncu_summary_cbox.json > /dev/null 2>> gpurun_out/${tag}_ncu_cbox.log
ncu --set full --clock-control none -k regex:aq_k_ -s 18 -c 17 -f -o gpurun_out/${tag}_prof_room \
    python tools/quick_bench.py --scene room --res 1920 1080 --spp 8 --reps 2 > gpurun_out/${tag}_ncu_room.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_prof_room.ncu-rep gpurun_out/${tag}_room_ncu_full.json > /dev/null 2>> gpurun_out/${tag}_ncu_room.log
python tools/ncu_wave_summary.py gpurun_out/${tag}_prof_room.ncu-rep --scene room --width 1920 --height 1080 --pool 16777216 \
    --source ${tag}_room_ncu_full.json --out gpurun_out/${tag}_ncu_summary_room.json > /dev/null 2>> gpurun_out/${tag}_ncu_room.log
rm -f gpurun_out/*.ncu-rep
if [ "$mode" != quick ]; then
  python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c2_ref.json 2>> gpurun_out/${tag}_bench.err
  python bench.py --scene room --width 1920 --height 1080 --spp 256 --cpu-spp 1 --strong-spp 0 > gpurun_out/${tag}_bench_c3.json 2>> gpurun_out/${tag}_bench.err
  ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
      python bench.py --steps 1 --warmup 1 --spp 128 --no-cpu-baseline --strong-spp 0 > gpurun_out/${tag}_ncu_launches.log 2>&1
else
  python bench.py --steps 2 --warmup 3 --spp 128 --cpu-spp 2 > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench.err
fi
for f in gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_stage.log gpurun_out/${tag}_bench.err; do tail -n 4 $f; done; du -sh gpurun_out
