#!/bin/bash
# Session-2 GPU call b: all GPU tests, nrc bench lines, ncu of the nrc kernels, refreshed C2 bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s2b_pytest.log
python tools/bench_nrc.py --scene room > gpurun_out/nrc_room.json 2> gpurun_out/nrc.err
python tools/bench_nrc.py --scene cbox --res 1024 1024 > gpurun_out/nrc_cbox.json 2>> gpurun_out/nrc.err
cat gpurun_out/nrc_room.json gpurun_out/nrc_cbox.json
ncu --set full --clock-control none --import-source on -k regex:aq_k_nrc_ -c 12 -f -o gpurun_out/prof_nrc \
    python tools/bench_nrc.py --scene room --quick > gpurun_out/ncu_nrc.log 2>&1
python bench.py > gpurun_out/bench_c2_s2.json 2> gpurun_out/bench_s2.err
tail -c 600 gpurun_out/bench_c2_s2.json
