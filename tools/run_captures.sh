#!/bin/bash
# Round capture set (run under gpurun, 1 GPU): parity tests, bench lines for C1/C2/C3, the
# CPU reference arm, C4 soup, ncu full captures of every kernel, ncu launch list of bench.py.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json
python bench.py --scene room --width 1920 --height 1080 --spp 256 --cpu-spp 1 > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
python bench.py --width 256 --height 256 --spp 16 --steps 5 --cpu-spp 16 > gpurun_out/bench_c1.json 2>> gpurun_out/bench.err
timeout 900 python tools/bench_soup.py > gpurun_out/soup_10m.json 2> gpurun_out/soup.err
ncu --set full --clock-control none --import-source on -k regex:aq_k_ -s 17 -c 8 -f -o gpurun_out/prof_cbox \
    python tools/quick_bench.py --spp 16 --reps 1 > gpurun_out/ncu_cbox.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aq_k_ -s 0 -c 8 -f -o gpurun_out/prof_room \
    python tools/quick_bench.py --scene room --res 1920 1080 --spp 4 --reps 1 > gpurun_out/ncu_room.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aq_k_trace -c 2 -f -o gpurun_out/prof_soup \
    python tools/bench_soup.py --tris 10000000 --rays 16777216 --reps 1 --check-brute 0 --check-bvh 0 > gpurun_out/ncu_soup.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --spp 128 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ls gpurun_out
